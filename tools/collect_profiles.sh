#!/bin/bash
# Round-end evidence run on one B200 (everything lands in gpurun_out/): GPU tests, bench lines, CUPTI table, phase
# breakdown, ncu launch list and ncu --set full captures of the three attention kernels + one GEMM.
mkdir -p gpurun_out
(timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log); tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/kernel_table.txt > /dev/null 2>&1
python bench.py --workload image --no-cpu-baseline > gpurun_out/bench_image.json 2> gpurun_out/bench_image.err
python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 200 python tools/phase_breakdown.py > gpurun_out/phases.txt 2>&1
timeout 120 python tools/bench_small.py > gpurun_out/small_kernels.txt 2>&1
timeout 120 python tools/bench_attn.py > gpurun_out/attn_shapes.txt 2>&1
timeout 120 python tools/bench_win.py > gpurun_out/win_shapes.txt 2>&1
timeout 200 python tools/bench_gemm.py > gpurun_out/gemm_shapes.txt 2>&1
T=96 NOBJ=13 timeout 300 python tools/run_multi_object.py > gpurun_out/multi_object_13.txt 2>&1
ONLY=mem.cross.dv.209k timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn_tc2 -s 2 -c 1 -f -o gpurun_out/attn2 python tools/bench_attn.py > /dev/null 2>&1
ONLY=hiera.global.b8 timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 2 -c 1 -f -o gpurun_out/attn96 python tools/bench_attn.py > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:win_attn_tc_kernel -s 22 -c 1 -f -o gpurun_out/win14 python tools/bench_win.py > /dev/null 2>&1
ONLY=s3.qkv timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -f -o gpurun_out/gemm_s3qkv python tools/bench_gemm.py > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv python bench.py --slices 24 --steps 1 --warmup 1 --no-graphs --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo collected
