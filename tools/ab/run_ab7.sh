tools/ab/lib_ab.sh "16" d: 3 main p1 p2 p3 > gpurun_out/af_ab7.jsonl
unset VLASOV_B200_LIB
python tools/ab/af_ab2.py 100000000 "1024" r8:af=1 r16:af=1,no_repg=1 >> gpurun_out/af_ab7.jsonl
python - <<EOF
import json
for l in open("gpurun_out/af_ab7.jsonl"):
    d=json.loads(l); print(d.get("lib","-"), d.get("round",0), d["variant"], d["n_h"], "%.4f %.4f frac %.3f" % (d["step_ms_median"], d["step_ms_min"], d["step_hbm_frac"]))
EOF
