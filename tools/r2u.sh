out=gpurun_out; mkdir -p $out
run() { # tag, env
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_n2_$1.json 2> $out/bench_n2_$1.err
  python - <<PY
import json
l = json.loads(open("$out/bench_n2_$1.json").read().strip().splitlines()[-1])
print("$1 N=2 value %.0f e2e %.0f pageable %.0f frac %.3f clocks %s" % (l["value"], l["e2e"]["value"], l["e2e"]["pageable"]["value"], l["roofline"]["frac"], l["clocks"]))
print("   per_rank", l["per_rank"])
c = l["c4"]; print("   c4 value %.0f ms_total %.2f per_rank %s" % (c["value"], c["ms_total"], c["per_rank_ms"]))
PY
}
run nonvml BENCH_NO_NVML=1 29514
run sampler BENCH_X=1 29515
run sampler2 BENCH_X=1 29516
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_n1_r2u.json 2> $out/bench_n1_r2u.err
python - <<PY
import json
l = json.loads(open("$out/bench_n1_r2u.json").read().strip().splitlines()[-1])
print("N=1 value %.0f e2e %.0f pageable %.0f frac %.3f clocks %s" % (l["value"], l["e2e"]["value"], l["e2e"]["pageable"]["value"], l["roofline"]["frac"], l["clocks"]))
c = l["c4"]; print("   c4 value %.0f ms_total %.2f" % (c["value"], c["ms_total"]))
PY
