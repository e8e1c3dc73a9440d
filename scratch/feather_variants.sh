#!/bin/bash
# builds libstitchb200_<name>.so variants of the streaming feather kernel shape (sb_fused.h SB_CFG_*) into scratch/variants/
set -e
cd "$(dirname "$0")/../stitchingvideo_b200/csrc"
OUT=../../scratch/variants; mkdir -p $OUT
build() {  # name H consumers producers ctas ringKB
  local d=/tmp/w/var_$1; mkdir -p $d
  local F="-DSB_CFG_FTT_H=$2 -DSB_CFG_FTS_CONSUMER_WARPS=$3 -DSB_CFG_FTS_PRODUCER_WARPS=$4 -DSB_CFG_FTS_CTAS_PER_SM=$5 -DSB_CFG_FTS_RING_KB=$6"
  local objs=""
  for f in capi_common capi_warper capi_comp capi_blender capi_compositor kernels_warp kernels_pointwise kernels_pyr kernels_blend kernels_fused kernels_feather_tma kernels_mb; do
    case $f in capi_compositor|kernels_feather_tma|kernels_fused)
      nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-ffp-contract=off --fmad=false $F -c $f.cu -o $d/$f.o & objs="$objs $d/$f.o";;
    *) objs="$objs $f.o";;
    esac
  done
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libstitchb200_$1.so $objs host_projector.o -cudart static
  echo built $1
}
build h32w16 32 16 4 2 100 &
build h16w8  16 8 2 3 64 &
build h32w8  32 8 2 3 64 &
build h16w16 16 16 4 2 100 &
build h32w16p2 32 16 2 2 100 &
build h64w16 64 16 4 2 100 &
wait
ls -la $OUT
