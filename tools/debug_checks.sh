#!/bin/bash
# Debug build of libuqoc.so with device-side bounds asserts and exchange epoch tags (-DUQOC_DEBUG_CHECKS, see
# csrc/uqoc_common.cuh), then the GPU parity tests on it.  The substitute for compute-sanitizer memcheck / racecheck on
# pools that refuse the tool:   tools/debug_checks.sh build   (here, no GPU)   /   tools/debug_checks.sh run   (GPU box)
set -e
cd "$(dirname "$0")/.."
PKG=universal_quantum_optimal_control_b200
OUT=$PKG/lib/variants/debug.so
if [ "$1" = "build" ]; then
  mkdir -p $PKG/lib/variants $PKG/build/debug
  FLAGS="-O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -DUQOC_DEBUG_CHECKS=1"
  for f in uqoc_api uqoc_su2_f32 uqoc_su2_f32_fast uqoc_su2_f64 uqoc_su4; do
    nvcc $FLAGS -c $PKG/csrc/$f.cu -o $PKG/build/debug/$f.o &
  done
  wait
  nvcc -shared -o $OUT $PKG/build/debug/*.o -gencode arch=compute_100a,code=sm_100a -cudart static
  echo built $OUT
else
  UQOC_LIB=$PWD/$OUT python -m pytest tests/test_gpu_su2.py tests/test_gpu_next_rows.py tests/test_gpu_pinned_rows.py tests/test_gpu_su4.py -x -q
fi
