out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_parity.py tests/test_scanmatcher.py tests/test_gpu_baseline_sizes.py -m gpu -q > $out/pytest_r2h.log 2>&1; tail -3 $out/pytest_r2h.log
timeout 300 python tools/diag_batch.py 20 > $out/diag_batch_r2h.log 2>&1; cat $out/diag_batch_r2h.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-c4 --no-cpu-baseline > $out/bench_r2h.json 2> $out/bench_r2h.err
B200REG_LIB_VARIANT=libb200reg_s4.so timeout 300 python bench.py --steps 20 --warmup 5 --slots 4 --no-c4 --no-cpu-baseline > $out/bench_s4_r2h.json 2> $out/bench_s4_r2h.err
B200REG_LIB_VARIANT=libb200reg_s4.so timeout 300 python bench.py --steps 20 --warmup 5 --slots 3 --no-c4 --no-cpu-baseline > $out/bench_s4as3_r2h.json 2> $out/bench_s4as3_r2h.err
python - <<PY
import json
for f in ["bench_r2h", "bench_s4_r2h", "bench_s4as3_r2h"]:
    try:
        l = json.loads(open("$out/" + f + ".json").read().strip().splitlines()[-1])
        print(f, "value %.0f  e2e %.0f  pageable %.0f  single %.0f  frac %.3f  us/eval %.2f" % (l["value"], l["e2e"]["value"], l["e2e"]["pageable"]["value"], l["single_align"]["value"], l["roofline"]["frac"], l["roofline"]["us_per_evaluation"]))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 $out/bench_s4_r2h.err
