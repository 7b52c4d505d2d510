#!/bin/bash
# Build kernel variants (different -D tuning macros) of libpkwhir.so into provekit_b200/variants/ for A/B runs on the GPU box:
#   tools/variants.sh name1 "-DPK_X=1 -DPK_Y=2" name2 "..."      then   PKWHIR_LIB=provekit_b200/variants/name1.so python tools/microbench.py
set -e
cd "$(dirname "$0")/../provekit_b200/csrc"
make -s
mkdir -p ../variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v $flags -c kernels.cu -o /tmp/kernels_$name.o 2> ../variants/$name.ptxas.log
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/$name.so /tmp/kernels_$name.o pkwhir.o host/*.o -lcudart -ldl
  echo "$name: $(grep -A2 ntt_pass ../variants/$name.ptxas.log | grep Used)"
done
