tools/ab/lib_ab.sh "16 32 64 256 1024" af:af=1 2 main w1 w2 w3 w4 w5 > gpurun_out/af_ab6.jsonl
unset VLASOV_B200_LIB
python tools/ab/af_ab2.py 100000000 "16 24 32 40" base:af=-1 af:af=1 >> gpurun_out/af_ab6.jsonl
python - <<EOF
import json
for l in open("gpurun_out/af_ab6.jsonl"):
    d=json.loads(l); print(d.get("lib","-"), d.get("round",0), d["variant"], d["n_h"], "%.4f %.4f frac %.3f" % (d["step_ms_median"], d["step_ms_min"], d["step_hbm_frac"]))
EOF
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
