#!/bin/bash
# Trimmed single-GPU verification + evidence bundle for the end of a round (gpurun --timeout 1200 -- 'bash tools/gpu_final.sh <tag>'):
# the gpu tests, smoke(), the default bench line, the reference arm, the launch list and one full capture of the solver kernel.
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q --durations=6 > $out/pytest_gpu_$tag.log 2>&1; tail -12 $out/pytest_gpu_$tag.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; tail -2 $out/smoke_$tag.log
timeout 400 python bench.py --impl reference --steps 5 --warmup 2 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err; tail -c 300 $out/bench_ref_$tag.json
timeout 600 python bench.py --steps 20 --warmup 5 > $out/bench_$tag.json 2> $out/bench_$tag.err; tail -5 $out/bench_$tag.err
python - <<PY
import json
for f in ["bench_$tag"]:
    try:
        l = json.loads(open("$out/" + f + ".json").read().strip().splitlines()[-1])
        print(f, "value %.0f  e2e %.0f  pageable %.0f  single %.0f  frac %.3f  us/eval %.2f cpu %.1f (%s thr) c4 %.0f" % (l["value"], l["e2e"]["value"], l["e2e"]["pageable"]["value"], l["single_align"]["value"], l["roofline"]["frac"], l["roofline"]["us_per_evaluation"], l["cpu_baseline"]["value"], l["cpu_baseline"]["cores"], l["c4"]["value"]))
        print("  clocks", l["clocks"])
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv \
  python tools/profile_step.py headline 3 20 > $out/prof_step_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ndt_solver_kernel -s 4 -c 1 \
  -o $out/prof_ndt_solver_$tag -f python tools/profile_step.py headline 3 20 > $out/prof_full_$tag.log 2>&1
ncu -i $out/prof_ndt_solver_$tag.ncu-rep --page raw --csv > $out/ndt_solver_raw_$tag.csv 2>/dev/null
ncu -i $out/prof_ndt_solver_$tag.ncu-rep --page details --csv > $out/ndt_solver_details_$tag.csv 2>/dev/null
ls -la $out | tail -4
