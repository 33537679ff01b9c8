#!/bin/bash
# round-2 measurement pass (one B200): bench line, reference arm, ncu launch list of one steady-state step,
# ncu --set full of the tensor-core kernels and of the BatchNorm streaming kernels (raw metrics exported as csv;
# the .ncu-rep files of two representative launches are kept for the source-level stall tables)
set -x
mkdir -p gpurun_out
T=${1:-r2}
python bench.py --profile-out gpurun_out/${T}_events.json > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 260 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --batch 1024 --steps 1 --warmup 3 --no-graph --no-cpu --no-sweep --no-ref-cuda --no-sustained > gpurun_out/${T}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"igemm_tma|igemm_pair|igemm_patch|wgrad_tma" -s 100 -c 56 -o /tmp/prof_gemm -f \
    python bench.py --batch 1024 --steps 1 --warmup 3 --no-graph --no-cpu --no-sweep --no-ref-cuda --no-sustained > gpurun_out/${T}_ncu_gemm.log 2>&1
ncu -i /tmp/prof_gemm.ncu-rep --page raw --csv > gpurun_out/${T}_gemm_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:"bn_|adam|peer" -s 100 -c 20 -o /tmp/prof_bn -f \
    python bench.py --batch 1024 --steps 1 --warmup 3 --no-graph --no-cpu --no-sweep --no-ref-cuda --no-sustained > gpurun_out/${T}_ncu_bn.log 2>&1
ncu -i /tmp/prof_bn.ncu-rep --page raw --csv > gpurun_out/${T}_bn_raw.csv 2>/dev/null
# source-level captures of the two kernels VERDICT r1 asked for (wide forward layer, N = 32-channel patch layer)
ncu --set full --clock-control none --import-source on -k regex:igemm_pair -s 3 -c 1 -o gpurun_out/${T}_src_deconv1fwd -f python tools/bench_layers.py deconv1.fwd > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:igemm_patch -s 3 -c 1 -o gpurun_out/${T}_src_deconv3fwd -f python tools/bench_layers.py deconv3.fwd > /dev/null 2>&1
python tools/bench_layers.py > gpurun_out/${T}_layers.txt 2>&1
python tools/bench_bn.py > gpurun_out/${T}_bn.txt 2>&1
python -m pytest tests -m gpu -q -s > gpurun_out/${T}_gpu_tests.log 2>&1; tail -3 gpurun_out/${T}_gpu_tests.log
ls -la gpurun_out | tail -15; du -sh gpurun_out
