out=gpurun_out; mkdir -p $out
B200REG_TRACE=1 timeout 300 python tools/diag_c4.py 4 > $out/diag_c4_trace_r2g.log 2>&1; grep -m12 "trace" $out/diag_c4_trace_r2g.log; tail -4 $out/diag_c4_trace_r2g.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $out/bench_n2_r2g.json 2> $out/bench_n2_r2g.err; tail -c 2500 $out/bench_n2_r2g.json; tail -5 $out/bench_n2_r2g.err
