out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_scanmatcher.py tests/test_gpu_deskew.py tests/test_gpu_baseline_sizes.py tests/test_gpu_gicp.py tests/test_gpu_parity.py -m gpu -q > $out/pytest_r2q.log 2>&1; tail -3 $out/pytest_r2q.log
timeout 600 python bench.py --workload c5 --frames 120 > $out/bench_c5_r2q.json 2> $out/bench_c5_r2q.err; python -c "
import json; l=json.loads(open('$out/bench_c5_r2q.json').read().strip().splitlines()[-1]); print('c5', l['value'], l['config']['passes_ms_per_frame'], l['cpu_baseline']['value'], l['cpu_baseline']['pose_parity_max_m'])"
tail -2 $out/bench_c5_r2q.err
