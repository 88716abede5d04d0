out=gpurun_out; mkdir -p $out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > $out/bench_n2_r2p.json 2> $out/bench_n2_r2p.err; tail -3 $out/bench_n2_r2p.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 5 --warmup 2 > $out/bench_ref_n2_r2p.json 2> $out/bench_ref_n2_r2p.err; tail -c 300 $out/bench_ref_n2_r2p.json
python - <<PY
import json
l = json.loads(open("$out/bench_n2_r2p.json").read().strip().splitlines()[-1])
print("N=2 value %.0f e2e %.0f pageable %.0f single %.0f frac %.3f per_rank %s" % (l["value"], l["e2e"]["value"], l["e2e"]["pageable"]["value"], l["single_align"]["value"], l["roofline"]["frac"], l["per_rank"]))
c = l["c4"]; print("c4 value %.0f ms_total %.2f per_rank %s cpu %s" % (c["value"], c["ms_total"], c["per_rank_ms"], c.get("cpu_baseline", {}).get("value")))
print("clocks", l["clocks"], "cpu", l["cpu_baseline"]["value"], l["cpu_baseline"]["cores"])
PY
