#!/bin/sh
# builds tools/_bin/gauss_dev (developer harness, see tools/gauss_dev.cu) against the in-tree library
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_bin
nvcc -std=c++17 -O2 -fmad=false -gencode arch=compute_100a,code=sm_100a -I include tools/gauss_dev.cu -o tools/_bin/gauss_dev \
    -L mathmap_b200 -lmathmap_b200 -Xlinker -rpath,'$ORIGIN/../../mathmap_b200'
