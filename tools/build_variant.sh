#!/bin/bash
# Builds a variant of the library with extra -D switches for kernel experiments: tools/build_variant.sh NAME [-DFOO ...]
# -> small-object-detection-transformers_b200/build/variants/libsodt_NAME.so (use with SODT_B200_LIB=...)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
out=small-object-detection-transformers_b200/build/variants
mkdir -p $out/obj_$name
for f in small-object-detection-transformers_b200/csrc/*.cu; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $f -o $out/obj_$name/$(basename $f .cu).o &
done
wait
/usr/local/cuda/bin/nvcc -shared -o $out/libsodt_$name.so $out/obj_$name/*.o -gencode arch=compute_100a,code=sm_100a
rm -rf $out/obj_$name
echo $out/libsodt_$name.so
