#!/bin/bash
# Builds libpeuqse_b200.so (sm_100a) in-tree. nvcc cross-compiles without a GPU.
set -e
cd "$(dirname "$0")/peuqse_b200/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -Xptxas -v"
mkdir -p ../_build
pids=()
for f in problem logp sweep sweep_grid mh mh_steps fullcov ess ess_rounds comm peer stats doe diag; do
  ( $NVCC $FLAGS -c $f.cu -o ../_build/$f.o > ../_build/$f.log 2>&1 || { cat ../_build/$f.log; exit 1; } ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o ../libpeuqse_b200.so ../_build/*.o -lcudart -ldl
echo "built peuqse_b200/libpeuqse_b200.so"
