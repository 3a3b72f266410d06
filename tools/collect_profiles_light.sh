mkdir -p gpurun_out
(MS2_PARITY_TABLE=gpurun_out/parity_table.jsonl timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log); tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity-check --no-strong --kernel-table gpurun_out/kernel_table.txt > /dev/null 2>&1
timeout 300 python bench.py --workload image --no-cpu-baseline > gpurun_out/bench_image.json 2> gpurun_out/bench_image.err
timeout 200 python tools/phase_breakdown.py > gpurun_out/phases.txt 2>&1
timeout 120 python tools/bench_attn.py > gpurun_out/attn_shapes.txt 2>&1
timeout 120 python tools/bench_win.py > gpurun_out/win_shapes.txt 2>&1
timeout 200 python tools/bench_gemm.py > gpurun_out/gemm_shapes.txt 2>&1
echo collected
