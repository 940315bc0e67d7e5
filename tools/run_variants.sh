#!/bin/bash
# time the variant builds of tools/variants.sh on one GPU:  tools/run_variants.sh "shape args" name1 name2 ...
args="$1"; shift
for v in "$@"; do
  UQOC_LIB=$PWD/universal_quantum_optimal_control_b200/lib/variants/$v.so timeout 120 python tools/shape_time.py $args 2>&1 | grep -v Warn
done
