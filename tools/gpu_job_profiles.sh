#!/bin/bash
# Round-2 evidence run on one B200: GPU tests, bench, launch lists, full ncu captures (each ncu pass only after the
# same command exited 0 without ncu).  Outputs land in gpurun_out/ (scratch); summaries are copied to profiles/.
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2b_pytest.log
python bench.py --steps 10 --warmup 3 > $O/r2b_bench_n1.json 2> $O/r2b_bench_n1.err || exit 1
python bench.py --no-cpu-baseline --steps 3 --warmup 3 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_bench_curriculum_launches.csv \
    python bench.py --no-cpu-baseline --steps 3 --warmup 3 > $O/ncu_l1.log 2>&1
python bench.py --no-cpu-baseline --workload grape --steps 3 --warmup 3 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_bench_grape_launches.csv \
    python bench.py --no-cpu-baseline --workload grape --steps 3 --warmup 3 > $O/ncu_l2.log 2>&1
python tools/profile_fwdbwd.py 4096 4096 256 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:su2_kernel -s 2 -c 1 -f -o $O/r2_prof_bench \
    python tools/profile_fwdbwd.py 4096 4096 256 > $O/ncu_p1.log 2>&1
python tools/profile_fwdbwd.py 1 65536 256 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:su2_kernel -s 2 -c 1 -f -o $O/r2_prof_grape \
    python tools/profile_fwdbwd.py 1 65536 256 > $O/ncu_p2.log 2>&1
ls -la $O | tail -20
