out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_pose_board.py -x -q 2>&1 | tail -5
run() { # tag, env
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-c4 > $out/bench_n2_$1.json 2> $out/bench_n2_$1.err
  python - <<PY
import json
try:
    l = json.loads(open("$out/bench_n2_$1.json").read().strip().splitlines()[-1])
    print("$1 N=2 value %.0f e2e %.0f pageable %.0f exchange %s checked %s" % (l["value"], l["e2e"]["value"], l["e2e"]["pageable"]["value"], l["details"]["pose_exchange"], l["details"]["pose_exchange_equals_nccl_all_gather"]))
    print("   per_rank", l["per_rank"])
except Exception as e:
    print("$1 failed", e); print(open("$out/bench_n2_$1.err").read()[-3000:])
PY
}
run board BENCH_X=1 29514
run nosync BENCH_SYNC_START=0 29515
run nccl BENCH_POSE_EXCHANGE=nccl 29516
