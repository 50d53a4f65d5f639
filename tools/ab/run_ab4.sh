python tools/ab/af_ab.py > gpurun_out/af_ab4.jsonl 2> gpurun_out/af_ab4.err
python - <<EOF
import json
for l in open("gpurun_out/af_ab4.jsonl"):
    d=json.loads(l); print(d["n_h"], d["variant"], "dep %.4f step %.4f frac %.3f err %s" % (d["deposit_ms"], d["step_ms"], d["step_hbm_frac"], d["rhs_rel_vs_base"]))
EOF
tail -3 gpurun_out/af_ab4.err; python -m pytest tests -m gpu -x -q 2>&1 | tail -4
