#!/bin/bash
# Per-role cycle counts of the role-parallel tick kernel (run under gpurun from the repo root):
#   gpurun --timeout 600 -- 'bash scripts/role_timing.sh c2 c3-compact c3-full c4'
# Builds an instrumented copy of the library (-DTL_ROLE_TIMING) next to the real one and runs a short bench with it.
set -u
mkdir -p gpurun_out
LIB=$PWD/townlet_b200/lib/libtownlet_b200_timing.so
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DTL_ROLE_TIMING \
  -Iinclude -o $LIB townlet_b200/csrc/townlet_b200.cu || exit 1
for W in "$@"; do
  echo "== $W"
  TOWNLET_B200_LIB=$LIB python bench.py --workload $W --no-cpu --no-e2e --steps 64 --warmup 32 --no-extra --min-launches 4 --min-seconds 0.001 2>&1 | grep TL_ROLE_TIMING | tail -9
done | tee gpurun_out/role_timing.txt
rm -f $LIB
