#!/bin/bash
# A/B of library builds on one box: tools/ab/lib_ab.sh "MESHES" "VARIANT-SPEC" ROUNDS name name ...   (tools/ab/lib_<name>.so)
M="$1"; V="$2"; R="$3"; shift 3
for r in $(seq 1 $R); do
  for L in "$@"; do
    if [ "$L" = "main" ]; then unset VLASOV_B200_LIB; else export VLASOV_B200_LIB=$PWD/tools/ab/lib_$L.so; fi
    python tools/ab/af_ab2.py 100000000 "$M" "$V" 2>/dev/null | sed "s/^{/{\"lib\": \"$L\", \"round\": $r, /"
  done
done
