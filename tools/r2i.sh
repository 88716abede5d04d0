out=gpurun_out; mkdir -p $out
run() { # tag variant slots
  B200REG_LIB_VARIANT=$2 timeout 300 python bench.py --steps 20 --warmup 5 --slots $3 --no-c4 --no-cpu-baseline > $out/bench_$1_r2i.json 2> $out/bench_$1_r2i.err
}
run base libb200reg.so 3
run p0s3 libb200reg_p0s3.so 3
run p0s4 libb200reg_p0s4.so 4
run p512s4 libb200reg_p512s4.so 4
run p768s4 libb200reg_p768s4.so 4
run p768s3 libb200reg_p768s3.so 3
run p0s4as3 libb200reg_p0s4.so 3
python - <<PY
import json
for f in ["base", "p0s3", "p0s4", "p512s4", "p768s4", "p768s3", "p0s4as3"]:
    try:
        l = json.loads(open("$out/bench_" + f + "_r2i.json").read().strip().splitlines()[-1])
        print(f, "value %.0f  e2e %.0f  single %.0f (kernel %.3f ms)  frac %.3f  us/eval %.2f" % (l["value"], l["e2e"]["value"], l["single_align"]["value"], l["single_align"]["kernel_ms_per_step"], l["roofline"]["frac"], l["roofline"]["us_per_evaluation"]))
    except Exception as e:
        print(f, "ERR", e)
PY
