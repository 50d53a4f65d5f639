#!/bin/bash
# A/B of bank-sorted pass builds (tools/ab/lib_<name>.so): step / deposit times at several mesh sizes, then warp
# instruction and shared-memory wavefront counts of one launch (ncu, two metrics only) at n_h = 256.
O=gpurun_out
for name in "$@"; do
  echo "== $name"
  VLASOV_B200_LIB=tools/ab/lib_$name.so python tools/ab/mesh_ab.py 100000000 256 512 1024 2>&1 | grep '"bankq": 1'
  VLASOV_B200_LIB=tools/ab/lib_$name.so ncu --metrics smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum --clock-control none -k regex:k_vp_pass_bq -s 4 -c 1 --csv python tools/ab/nh_run.py 100000000 256 0 0 2>/dev/null | tail -4 | cut -d, -f5,13-
done
