#!/bin/bash
# Build A/B variants of the library here (no GPU needed) so that one gpurun call can time them all:
#   tools/ab_variants.sh name1 "-DGHB_X=1" name2 "-DGHB_Y=0 -DGHB_Z=2" ...   ->  tools/_bin/libghb_<name>.so
# then on the GPU box:  GHB_LIB_PATH=tools/_bin/libghb_<name>.so python tools/ab_ll.py
cd "$(dirname "$0")/.." || exit 1
mkdir -p tools/_bin
while [ $# -ge 2 ]; do
  name="$1"; flags="$2"; shift 2
  (nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared $flags \
     -o tools/_bin/libghb_$name.so gridaphybrid.jl_b200/csrc/*.cu && echo "built $name ($flags)") &
done
wait
