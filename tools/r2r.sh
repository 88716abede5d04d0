out=gpurun_out; mkdir -p $out
for k in 1 2; do
timeout 600 python bench.py --workload c5 --frames 120 > $out/bench_c5_r2r$k.json 2> $out/bench_c5_r2r$k.err; python -c "
import json; l=json.loads(open('$out/bench_c5_r2r$k.json').read().strip().splitlines()[-1]); print('c5', l['value'], l['config']['passes_ms_per_frame'])"
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_r2r.json 2> $out/bench_r2r.err; python -c "
import json; l=json.loads(open('$out/bench_r2r.json').read().strip().splitlines()[-1]); print('headline', l['value'], l['e2e']['value'], l['e2e']['pageable']['value'], l['roofline']['frac'], l['c4']['ms_per_pair'], l['target_build'])"
