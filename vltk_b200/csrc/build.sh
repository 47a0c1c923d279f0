#!/bin/bash
# Builds libvltk_frcnn.so (sm_100a only) in-tree.  Usage: build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
# VLTK_TRACE=1 builds the diagnosis variant ../libvltk_frcnn_trace.so (pipeline trace hooks in conv_tc.cu compiled in)
OUT=../libvltk_frcnn.so
OBJ=_obj
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC $ARCH $@"
if [ "${VLTK_TRACE:-0}" = "1" ]; then OUT=../libvltk_frcnn_trace.so; OBJ=_obj_trace; COMMON="$COMMON -DVLTK_TC_TRACE"; fi
mkdir -p $OBJ
pids=()
for f in conv_simt conv_tc conv_tcx pack engine; do
  nvcc $COMMON -c $f.cu -o $OBJ/$f.o & pids+=($!)
done
# box arithmetic must round like the reference's unfused torch ops -> no FMA contraction
for f in elementwise h2ops rpn roipool tail jpeg; do
  nvcc $COMMON -fmad=false -c $f.cu -o $OBJ/$f.o & pids+=($!)
done
g++ -O3 -std=c++17 -fPIC -pthread -c jpeg_host.cpp -o $OBJ/jpeg_host.o & pids+=($!)
for p in "${pids[@]}"; do wait $p; done
nvcc -shared $ARCH -o $OUT $OBJ/*.o -cudart static
echo "built $OUT"
