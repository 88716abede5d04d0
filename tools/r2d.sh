out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_deskew.py tests/test_scanmatcher.py -m gpu -q > $out/pytest_r2d.log 2>&1; tail -4 $out/pytest_r2d.log
timeout 600 python bench.py --workload c5 --frames 120 > $out/bench_c5_r2d.json 2> $out/bench_c5_r2d.err; tail -c 900 $out/bench_c5_r2d.json; tail -3 $out/bench_c5_r2d.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_c4_r2d.csv python tools/profile_c4.py > $out/prof_c4_r2d.log 2>&1; tail -3 $out/prof_c4_r2d.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/launches_c3_r2d.csv python tools/profile_gicp.py 1 > $out/prof_c3_r2d.log 2>&1; tail -3 $out/prof_c3_r2d.log
