#!/bin/bash
# Evidence run on one B200: launch lists of the bench commands and full ncu captures of the top kernels (each ncu pass
# only after the same command exited 0 without ncu).  Outputs land in gpurun_out/ (scratch); tools/ncu_summary.py
# condenses the reports into profiles/.
set -x
O=gpurun_out
python bench.py --no-cpu-baseline --steps 3 --warmup 3 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_bench_curriculum_launches.csv \
    python bench.py --no-cpu-baseline --steps 3 --warmup 3 > $O/ncu_l1.log 2>&1
python bench.py --no-cpu-baseline --workload grape --steps 3 --warmup 3 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_bench_grape_launches.csv \
    python bench.py --no-cpu-baseline --workload grape --steps 3 --warmup 3 > $O/ncu_l2.log 2>&1
python tools/profile_fwdbwd.py 4096 4096 256 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:su2_kernel -s 2 -c 1 -f -o $O/r2_prof_bench_final \
    python tools/profile_fwdbwd.py 4096 4096 256 > $O/ncu_p1.log 2>&1
python tools/profile_fwdbwd.py 1 65536 256 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:su2_kernel -s 2 -c 1 -f -o $O/r2_prof_grape_final \
    python tools/profile_fwdbwd.py 1 65536 256 > $O/ncu_p2.log 2>&1
python tools/su4_probe.py 8 128 32768 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:su4e_kernel -s 14 -c 1 -f -o $O/r2_prof_su4e_final \
    python tools/su4_probe.py 8 128 32768 > $O/ncu_p3.log 2>&1
ls -la $O | tail -8
