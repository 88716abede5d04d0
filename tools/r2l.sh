out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_batch.py tests/test_gpu_gicp.py -m gpu -q > $out/pytest_r2l.log 2>&1; tail -3 $out/pytest_r2l.log
timeout 300 python tools/diag_batch.py 20 > $out/diag_batch_r2l.log 2>&1; cat $out/diag_batch_r2l.log
B200REG_TRACE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-c4 --no-cpu-baseline > $out/bench_trace_r2l.json 2> $out/bench_trace_r2l.err; grep "batch of 20" $out/bench_trace_r2l.err | tail -6
for e in 2 3 4; do
  B200REG_SWEEP_ENGINES=$e timeout 300 python tools/diag_c4.py 8 > $out/diag_c4_e${e}_r2l.log 2>&1; echo "engines $e"; tail -2 $out/diag_c4_e${e}_r2l.log
done
