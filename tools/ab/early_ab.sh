#!/bin/bash
# fixed cost of a fused step with / without the early particle loads of the dependent launch (one GPU)
for lib in vlasovmethods.jl_b200/libvlasov_b200.so tools/ab/lib_noearly.so; do
  echo "== $lib"
  VLASOV_B200_LIB=$lib python tools/ab/small_n.py
done
