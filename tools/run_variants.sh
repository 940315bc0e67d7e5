for v in r1 f0b0 f1b0 f0b1 f1b1 f0b2 f1b2; do
  UQOC_LIB=$PWD/universal_quantum_optimal_control_b200/lib/variants/$v.so timeout 100 python tools/shape_time.py flags=0x40000000 c5 c3 score 2>&1 | grep -v Warn
done
timeout 100 python tools/shape_time.py c3 c1 2>&1
