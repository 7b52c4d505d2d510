#!/bin/bash
# Build kernel variants (different -D tuning macros) of libpkwhir.so into provekit_b200/variants/ for A/B runs on the GPU box:
#   tools/variants.sh name1 "-DPK_X=1 -DPK_Y=2" name2 "..."      then   PKWHIR_LIB=provekit_b200/variants/name1.so python tools/microbench.py
# The flags reach every device translation unit (kernels.cu, glue.cu, ntt.cu); host objects are shared with the main build.
set -e
cd "$(dirname "$0")/../provekit_b200/csrc"
make -s
mkdir -p ../variants
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  : > ../variants/$name.ptxas.log
  for u in kernels glue ntt; do
    $NV $flags -c $u.cu -o /tmp/${u}_$name.o 2>> ../variants/$name.ptxas.log &
  done
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/$name.so /tmp/kernels_$name.o /tmp/glue_$name.o /tmp/ntt_$name.o pkwhir.o host/*.o -lcudart -ldl
  echo "$name: built; kernels with spills: $(grep 'spill stores' ../variants/$name.ptxas.log | grep -vc ' 0 bytes spill stores')"
done
