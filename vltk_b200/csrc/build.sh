#!/bin/bash
# Builds libvltk_frcnn.so (sm_100a only) in-tree.  Usage: build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
OUT=../libvltk_frcnn.so
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC $ARCH $@"
mkdir -p _obj
pids=()
for f in conv_simt conv_tc pack engine; do
  nvcc $COMMON -c $f.cu -o _obj/$f.o & pids+=($!)
done
# box arithmetic must round like the reference's unfused torch ops -> no FMA contraction
for f in elementwise rpn roipool tail jpeg; do
  nvcc $COMMON -fmad=false -c $f.cu -o _obj/$f.o & pids+=($!)
done
g++ -O3 -std=c++17 -fPIC -pthread -c jpeg_host.cpp -o _obj/jpeg_host.o & pids+=($!)
for p in "${pids[@]}"; do wait $p; done
nvcc -shared $ARCH -o $OUT _obj/*.o -cudart static
echo "built $OUT"
