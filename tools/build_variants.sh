#!/bin/bash
# Tuning aid: builds libfealpy_b200 variants that differ in compile-time knobs of assemble.cu / cg.cu into
# fealpy_b200/csrc/build/variants/<name>.so (git-ignored, travels to the GPU box); select one with FB2_LIB_PATH.
#   tools/build_variants.sh name "-DFB2_ASM4_WARPS=3 -DFB2_ASM4_MINBLOCKS=3" [file.cu ...]
set -e
cd "$(dirname "$0")/../fealpy_b200/csrc"
name=$1; flags=$2; shift 2
files=${@:-assemble.cu}
mkdir -p build/variants/$name
ARCH="-gencode arch=compute_100a,code=sm_100a"
objs=""
for f in capi.cu sort_scan.cu elem.cu coo_csr.cu cg.cu topo.cu assemble.cu bc_source.cu peer.cu; do
  if [[ " $files " == *" $f "* ]]; then
    nvcc $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O3 $flags -c $f -o build/variants/$name/${f%.cu}.o
    objs="$objs build/variants/$name/${f%.cu}.o"
  else
    objs="$objs build/${f%.cu}.o"
  fi
done
nvcc $ARCH -shared -o build/variants/$name.so $objs -cudart static
echo "built build/variants/$name.so"
