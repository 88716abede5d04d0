out=gpurun_out; mkdir -p $out
for v in 250 1000 4000 none; do
  if [ $v = none ]; then export BENCH_NO_NVML=1; else export BENCH_CLK_PERIOD_US=$v; fi
  for rep in 1 2; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-c4 > $out/bench_clk${v}_$rep.json 2> $out/bench_clk${v}_$rep.err; python -c "
import json; l=json.loads(open('$out/bench_clk${v}_$rep.json').read().strip().splitlines()[-1]); print('clk $v', round(l['value']), round(l['e2e']['value']), round(1e3*l['roofline']['launch_ms']), 'us kernel', l['clocks']['samples'], 'samples')"
  done
done
