tools/ab/lib_ab.sh "64 256 1024" af:af=1 2 v0 v1 v2 v3 > gpurun_out/af_ab5.jsonl
python - <<EOF
import json
for l in open("gpurun_out/af_ab5.jsonl"):
    d=json.loads(l); print(d["lib"], d["round"], d["n_h"], "%.4f %.4f frac %.3f" % (d["step_ms_median"], d["step_ms_min"], d["step_hbm_frac"]))
EOF
for n in 16 64 256 1024; do python tools/ab/burst.py 100000000 $n; sleep 5; done > gpurun_out/burst1.jsonl 2>&1
python tools/ab/burst.py 100000000 16 af=1 >> gpurun_out/burst1.jsonl; python tools/ab/burst.py 100000000 64 af=-1 >> gpurun_out/burst1.jsonl
cat gpurun_out/burst1.jsonl
tools/ab/ncu_af.sh > gpurun_out/r02c_ncu_af_digest.txt 2>&1; cat gpurun_out/r02c_ncu_af_digest.txt
