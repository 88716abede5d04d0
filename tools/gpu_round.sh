#!/bin/bash
# Standard single-GPU verification + evidence bundle, run under gpurun from the repo root:
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag>'
# Outputs land in gpurun_out/ (copy what should be judged into profiles/).
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > $out/pytest_gpu_$tag.log 2>&1; tail -14 $out/pytest_gpu_$tag.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; tail -2 $out/smoke_$tag.log
timeout 600 python bench.py --steps 20 --warmup 5 > $out/bench_$tag.json 2> $out/bench_$tag.err; tail -c 1500 $out/bench_$tag.json; tail -5 $out/bench_$tag.err
for s in 2 1; do
  timeout 300 python bench.py --steps 20 --warmup 5 --slots $s --no-c4 --no-cpu-baseline > $out/bench_slots${s}_$tag.json 2> $out/bench_slots${s}_$tag.err
done
for w in c2 c1; do
  timeout 300 python bench.py --steps 20 --warmup 5 --workload $w --no-c4 > $out/bench_${w}_$tag.json 2> $out/bench_${w}_$tag.err
done
python - <<PY
import json
for f in ["bench_$tag", "bench_slots2_$tag", "bench_slots1_$tag", "bench_c2_$tag", "bench_c1_$tag"]:
    try:
        l = json.loads(open("$out/" + f + ".json").read().strip().splitlines()[-1])
        print(f, "value %.0f  e2e %.0f  pageable %.0f  single %.0f  frac %.3f  us/eval %.2f" % (l["value"], l["e2e"]["value"], l["e2e"]["pageable"]["value"], l["single_align"]["value"], l["roofline"]["frac"], l["roofline"]["us_per_evaluation"]))
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 400 python bench.py --impl reference --steps 5 --warmup 2 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err; tail -c 300 $out/bench_ref_$tag.json
timeout 300 python tools/diag_c4.py 8 > $out/diag_c4_$tag.log 2>&1; tail -6 $out/diag_c4_$tag.log
timeout 300 python tools/diag_batch.py 20 > $out/diag_batch_$tag.log 2>&1; cat $out/diag_batch_$tag.log
timeout 600 python bench.py --workload c3 > $out/bench_c3_$tag.json 2> $out/bench_c3_$tag.err; tail -c 1200 $out/bench_c3_$tag.json; tail -3 $out/bench_c3_$tag.err
timeout 600 python bench.py --workload c5 --frames 120 > $out/bench_c5_$tag.json 2> $out/bench_c5_$tag.err; tail -c 700 $out/bench_c5_$tag.json; tail -3 $out/bench_c5_$tag.err
# launch list of the profile command (cold-cache, serialised: compare shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv \
  python tools/profile_step.py headline 3 20 > $out/prof_step_$tag.log 2>&1
# one full capture of the solver kernel: launch #5 = the second HBM-resident batched launch of 20 registrations (3 single + 3 batched launches)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ndt_solver_kernel -s 4 -c 1 \
  -o $out/prof_ndt_solver_$tag -f python tools/profile_step.py headline 3 20 > $out/prof_full_$tag.log 2>&1
ncu -i $out/prof_ndt_solver_$tag.ncu-rep --page raw --csv > $out/ndt_solver_raw_$tag.csv 2>/dev/null
ncu -i $out/prof_ndt_solver_$tag.ncu-rep --page details --csv > $out/ndt_solver_details_$tag.csv 2>/dev/null
bash tools/profile_kernels.sh $tag > $out/profile_kernels_$tag.log 2>&1; tail -16 $out/profile_kernels_$tag.log
ls -la $out | tail -4
