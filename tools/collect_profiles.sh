#!/bin/bash
# Evidence run on one B200 (everything lands in gpurun_out/): GPU tests with the measured parity table, the default
# bench line, CUPTI kernel table, phase breakdown, per-shape timings, the ncu launch list of the bench command and
# ncu --set full captures of the dominant kernels.  Summaries are then copied into profiles/ (tools/ncu_hot.py).
mkdir -p gpurun_out
(MS2_PARITY_TABLE=gpurun_out/parity_table.jsonl timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log); tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check --no-strong --kernel-table gpurun_out/kernel_table.txt > /dev/null 2>&1
timeout 300 python bench.py --workload image --no-cpu-baseline > gpurun_out/bench_image.json 2> gpurun_out/bench_image.err
timeout 300 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 200 python tools/phase_breakdown.py > gpurun_out/phases.txt 2>&1
timeout 120 python tools/bench_attn.py > gpurun_out/attn_shapes.txt 2>&1
timeout 120 python tools/bench_win.py > gpurun_out/win_shapes.txt 2>&1
timeout 200 python tools/bench_gemm.py > gpurun_out/gemm_shapes.txt 2>&1
ONLY=hiera.global.b8 timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 2 -c 1 -f -o gpurun_out/attn96 python tools/bench_attn.py > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:win_attn_tc_kernel -s 22 -c 1 -f -o gpurun_out/win14 python tools/bench_win.py > /dev/null 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches.csv python bench.py --slices 24 --steps 1 --warmup 1 --no-graphs --no-cpu-baseline --no-gpu-baseline --no-parity-check --no-strong > gpurun_out/ncu_launch.log 2>&1
echo collected
