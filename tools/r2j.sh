out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_parity.py -m gpu -q > $out/pytest_r2j.log 2>&1; tail -3 $out/pytest_r2j.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-c4 --no-cpu-baseline > $out/bench_r2j.json 2> $out/bench_r2j.err
B200REG_NO_STREAM_MEMOPS=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-c4 --no-cpu-baseline > $out/bench_nomemops_r2j.json 2> $out/bench_nomemops_r2j.err
python - <<PY
import json
for f in ["bench_r2j", "bench_nomemops_r2j"]:
    try:
        l = json.loads(open("$out/" + f + ".json").read().strip().splitlines()[-1])
        print(f, "value %.0f  e2e %.0f  pageable %.0f  single %.0f  frac %.3f  us/eval %.2f bitwise %s" % (l["value"], l["e2e"]["value"], l["e2e"]["pageable"]["value"], l["single_align"]["value"], l["roofline"]["frac"], l["roofline"]["us_per_evaluation"], l["details"]["batch_bitwise_equals_single_align"]))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 $out/bench_r2j.err
