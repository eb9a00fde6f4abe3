#!/bin/bash
# Build libdmip_sm100.so in-tree (sm_100a only; nvcc cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
OUT=../libdmip_sm100.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v"
mkdir -p build
pids=()
for f in dmip_api dmip_pack dmip_tc dmip_f32 dmip_loss; do
  ( $NVCC $FLAGS -c $f.cu -o build/$f.o > build/$f.log 2>&1 || { cat build/$f.log; exit 1; } ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT build/dmip_api.o build/dmip_pack.o build/dmip_tc.o build/dmip_f32.o build/dmip_loss.o
echo "built $(realpath $OUT)"
