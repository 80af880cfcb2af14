#!/bin/bash
# Builds a tuning variant of the library: tools/build_variant.sh NAME -DFMHR_LB_SHADE=3 ...  -> gpurun_variants/libfmhr_NAME.so
set -e
NAME=$1; shift
mkdir -p /root/repo/variants/obj_$NAME
cd /root/repo/fmhr_b200/csrc
for f in api raster interpolate antialias mesh meshlet ncc_loop ham; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $f.cu -o /root/repo/variants/obj_$NAME/$f.o &
done
wait
nvcc -shared -o /root/repo/variants/libfmhr_$NAME.so /root/repo/variants/obj_$NAME/*.o -lcudart
rm -rf /root/repo/variants/obj_$NAME
echo built variants/libfmhr_$NAME.so
