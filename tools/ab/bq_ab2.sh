#!/bin/bash
for name in "$@"; do
  echo "== $name"
  VLASOV_B200_LIB=tools/ab/lib_$name.so python tools/ab/mesh_ab.py 100000000 200 256 512 1024 2048 2>&1 | grep '"bankq": 1' | cut -c1-150
done
