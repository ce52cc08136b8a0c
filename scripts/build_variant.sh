#!/bin/bash
# Build an A/B variant of the library with extra nvcc defines:
#   bash scripts/build_variant.sh nacc6 -DDPC_RING_NACC=6
# -> pytorch-unsup-pc_b200/lib/variants/libdpc_b200_<name>.so  (use with DPC_B200_LIB=...)
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/pytorch-unsup-pc_b200/csrc
out=$root/pytorch-unsup-pc_b200/lib/variants
obj=$root/pytorch-unsup-pc_b200/build/variant_$name
mkdir -p $out $obj
flags="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr"
for f in api pose_scatter blur_xy drc scatter_sorted candidate_loss chamfer replica feature microbench; do
  extra=""; [ $f = chamfer ] && extra="-fmad=false"   # see csrc/Makefile
  nvcc $flags $extra "$@" -c $src/$f.cu -o $obj/$f.o &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $out/libdpc_b200_$name.so $obj/*.o -lcudart
echo built $out/libdpc_b200_$name.so
