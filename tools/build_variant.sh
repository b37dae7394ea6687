#!/bin/bash
# tools/build_variant.sh NAME "-DFOO=1 ..." : builds build/variants/NAME.so with extra nvcc defines
# (kernel-variant experiments; select at run time with CSB_LIB_PATH=build/variants/NAME.so).
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants/obj_$name
for f in cusift_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v -cudart shared \
       -Iinclude -Iinclude/cusift -Icusift_b200/csrc $@ -c $f -o build/variants/obj_$name/$b.o 2> build/variants/obj_$name/$b.log &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart shared -o build/variants/$name.so build/variants/obj_$name/*.o
grep -A1 "k_orient_desc\|k_find_points\|k_blur_dog2" build/variants/obj_$name/*.log | grep -E "registers|spill" | head -12
cp build/variants/obj_$name/kernels_pyramid.log build/variants/$name.pyramid.log; rm -rf build/variants/obj_$name
