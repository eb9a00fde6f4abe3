#!/bin/bash
# Build tools/probe/libdmip_probe.so: tcgen05 building-block self-tests and micro-benchmarks (sm_100a).  Test / tooling
# infrastructure only — the product library (csrc/build.sh) contains none of this.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -o libdmip_probe.so dmip_probe.cu
echo "built $(realpath libdmip_probe.so)"
