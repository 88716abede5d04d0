N=$1; out=gpurun_out; mkdir -p $out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_n${N}_final.json 2> $out/bench_n${N}_final.err
python - <<PY
import json
try:
    l = json.loads(open("$out/bench_n${N}_final.json").read().strip().splitlines()[-1])
    print("N=$N value %.0f e2e %.0f pageable %.0f exchange %s checked %s launches %s" % (l["value"], l["e2e"]["value"], l["e2e"]["pageable"]["value"], l["details"]["pose_exchange"], l["details"]["pose_exchange_equals_nccl_all_gather"], l["gpu_launches"]))
    for p in l["per_rank"]: print("   ", {k: (round(v, 4) if not isinstance(v, dict) else {a: round(b, 4) for a, b in v.items()}) for k, v in p.items()})
    c = l.get("c4")
    if c: print("   c4 value %.0f ms_total %.2f per_rank %s" % (c["value"], c["ms_total"], [round(x, 2) for x in c["per_rank_ms"]]))
    print("   clocks", l["clocks"])
except Exception as e:
    print("failed", e); print(open("$out/bench_n${N}_final.err").read()[-3000:])
PY
