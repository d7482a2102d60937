#!/bin/bash
# builds the measurement scaffolds under tools/probe/ into tools/_bin/ (git-ignored, travels to the GPU box)
set -e
cd "$(dirname "$0")/../.."
CCCL=${MATX_CCCL:-/opt/prime-rl/.venv/lib/python3.12/site-packages/flashinfer/data/cccl}
mkdir -p tools/_bin
for f in "$@"; do
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -w \
    -I$CCCL/libcudacxx/include -I$CCCL/cub -I$CCCL/thrust tools/probe/$f.cu -o tools/_bin/$f
done
