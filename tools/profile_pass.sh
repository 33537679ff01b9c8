#!/bin/bash
# round-1 final measurement pass
set -x
mkdir -p gpurun_out
python bench.py --profile-out gpurun_out/prof_r1f_b1024.json > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err
python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/bench_r1f_ref.json 2>> gpurun_out/bench_r1f.err
# launch list of ONE steady-state step (skip the allocation step and two more)
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 260 --csv --log-file gpurun_out/launches_r1f.csv \
    python bench.py --batch 1024 --steps 1 --warmup 3 --no-graph --no-cpu --no-sweep > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"igemm_tma|wgrad_tma" -s 120 -c 40 -o /tmp/prof_gemm_r1f -f \
    python bench.py --batch 1024 --steps 1 --warmup 3 --no-graph --no-cpu --no-sweep > gpurun_out/ncu_gemm.log 2>&1
ncu -i /tmp/prof_gemm_r1f.ncu-rep --page raw --csv > gpurun_out/prof_gemm_r1f_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:"bn_" -s 120 -c 16 -o /tmp/prof_bn_r1f -f \
    python bench.py --batch 1024 --steps 1 --warmup 3 --no-graph --no-cpu --no-sweep > gpurun_out/ncu_bn.log 2>&1
ncu -i /tmp/prof_bn_r1f.ncu-rep --page raw --csv > gpurun_out/prof_bn_r1f_raw.csv 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
ls -la gpurun_out | tail -12; du -sh gpurun_out
