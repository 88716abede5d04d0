out=gpurun_out; mkdir -p $out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_n2_r2s.json 2> $out/bench_n2_r2s.err
python - <<PY
import json
l = json.loads(open("$out/bench_n2_r2s.json").read().strip().splitlines()[-1])
print("N=2 value %.0f e2e %.0f pageable %.0f frac %.3f per_rank %s" % (l["value"], l["e2e"]["value"], l["e2e"]["pageable"]["value"], l["roofline"]["frac"], l["per_rank"]))
c = l["c4"]; print("c4 value %.0f ms_total %.2f per_rank %s" % (c["value"], c["ms_total"], c["per_rank_ms"]))
PY
tail -3 $out/bench_n2_r2s.err
