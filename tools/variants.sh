#!/bin/bash
# Build tuning variants of libuqoc.so (different -D macros for the packed SU(2) kernel) into
# universal_quantum_optimal_control_b200/lib/variants/<name>.so; select one at run time with UQOC_LIB=<path>.
#   tools/variants.sh name1:"-DX=1 -DY=2" name2:"-DZ=3" ...
# uqoc_api.cu is rebuilt too: the host-side launch plan sizes the shared memory from the same macros.
set -e
cd "$(dirname "$0")/../universal_quantum_optimal_control_b200"
(cd .. && python -m universal_quantum_optimal_control_b200.build >/dev/null)
mkdir -p lib/variants build/variants
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr"
for spec in "$@"; do
  name="${spec%%:*}"; defs="${spec#*:}"
  (
    nvcc $FLAGS $defs -c csrc/uqoc_su2_f32.cu -o build/variants/$name.o &
    nvcc $FLAGS $defs -c csrc/uqoc_api.cu -o build/variants/$name.api.o &
    wait
    nvcc -shared -o lib/variants/$name.so build/variants/$name.api.o build/variants/$name.o build/uqoc_su2_f32_fast.o build/uqoc_su2_f64.o build/uqoc_su4.o -gencode arch=compute_100a,code=sm_100a -cudart static
    echo built $name
  ) &
done
wait
