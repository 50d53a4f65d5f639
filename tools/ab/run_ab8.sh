python tools/ab/af_ab2.py 100000000 "256 512 1024" d: norepg:no_repg=1 r16:af_replicas=16 r16norepg:af_replicas=16,no_repg=1 nopdl:no_pdl=1 > gpurun_out/af_ab8.jsonl
python - <<EOF
import json
for l in open("gpurun_out/af_ab8.jsonl"):
    d=json.loads(l); print(d["variant"], d["n_h"], "%.4f %.4f frac %.3f" % (d["step_ms_median"], d["step_ms_min"], d["step_hbm_frac"]))
EOF
