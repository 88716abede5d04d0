out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_parity.py tests/test_gpu_gicp.py tests/test_gpu_baseline_sizes.py -m gpu -q > $out/pytest_r2e.log 2>&1; tail -5 $out/pytest_r2e.log
timeout 300 python tools/diag_c4.py 8 > $out/diag_c4_r2e.log 2>&1; tail -6 $out/diag_c4_r2e.log
timeout 600 python bench.py --workload c3 --no-cpu-baseline > $out/bench_c3_r2e.json 2> $out/bench_c3_r2e.err; tail -c 500 $out/bench_c3_r2e.json
timeout 300 python bench.py --workload c4 > $out/bench_c4_r2e.json 2> $out/bench_c4_r2e.err; tail -c 600 $out/bench_c4_r2e.json; tail -3 $out/bench_c4_r2e.err
