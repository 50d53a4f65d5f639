#!/bin/bash
# Build an A/B variant of the library: only the order-4 pass unit is recompiled (with the given extra nvcc flags) and
# linked with the objects of the last full build:  tools/ab/build_variant.sh NAME "-DVM_X=1 ..."
# -> tools/ab/lib_NAME.so (git-ignored, travels with the gpurun snapshot; selected with VLASOV_B200_LIB)
set -e
cd "$(dirname "$0")/../../vlasovmethods.jl_b200/csrc"
mkdir -p /tmp/abbuild/$1
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a"
$NV -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ -I../../include -I. --expt-relaxed-constexpr $2 \
    -DVM_PASS_ORDER=4 -c vm_pass_order.cu -o /tmp/abbuild/$1/vm_pass_k4.o
$NV -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ -I../../include -I. --expt-relaxed-constexpr $2 \
    -c vm_push.cu -o /tmp/abbuild/$1/vm_push.o          # (the host-side plan lives in the headers too)
OBJS=$(ls build/*.o | grep -v -e vm_pass_k4 -e vm_push)
$NV -shared -ccbin /usr/bin/g++ -o ../../tools/ab/lib_$1.so $OBJS /tmp/abbuild/$1/vm_pass_k4.o /tmp/abbuild/$1/vm_push.o -ldl
echo built tools/ab/lib_$1.so
