#!/bin/bash
# build_variant.sh NAME "-DSB_CFG_...=.. ..." : a tuning variant of the library as variants/libstitchb200_NAME.so
# (select it with STITCHB200_LIB=...; variants/ is git-ignored but travels to the GPU box)
set -e
NAME=$1; DEFS=$2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/stitchingvideo_b200/csrc
OUT=$ROOT/variants; mkdir -p $OUT/obj_$NAME
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-ffp-contract=off --fmad=false $DEFS"
# only the files that see the tuning macros are rebuilt; the rest are taken from the main build
for f in kernels_fstream2 capi_compositor kernels_mb kernels_mb_pyr kernels_mb_stream kernels_feather_tma; do
  $NV -c $SRC/$f.cu -o $OUT/obj_$NAME/$f.o 2> $OUT/obj_$NAME/$f.log &
done
wait
OBJS=""
for o in $SRC/*.o; do b=$(basename $o); if [ -f $OUT/obj_$NAME/$b ]; then OBJS="$OBJS $OUT/obj_$NAME/$b"; else OBJS="$OBJS $o"; fi; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libstitchb200_$NAME.so $OBJS -cudart static
echo built $OUT/libstitchb200_$NAME.so
