out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_gicp.py tests/test_gpu_parity.py -m gpu -q > $out/pytest_r2m.log 2>&1; tail -3 $out/pytest_r2m.log
B200REG_GICP_TRACE=1 timeout 300 python tools/profile_gicp.py 2 > $out/gicp_trace_r2m.log 2>&1; grep "trace\|align" $out/gicp_trace_r2m.log | tail -5
B200REG_GICP_COV_SCALAR=1 B200REG_GICP_TRACE=1 timeout 300 python tools/profile_gicp.py 2 > $out/gicp_trace_scalar_r2m.log 2>&1; grep "trace\|align" $out/gicp_trace_scalar_r2m.log | tail -3
timeout 600 python bench.py --workload c3 --no-cpu-baseline > $out/bench_c3_r2m.json 2> $out/bench_c3_r2m.err; python -c "
import json; l=json.loads(open('$out/bench_c3_r2m.json').read().strip().splitlines()[-1]); print('c3', l['value'], l['ms_per_step'], l['roofline']['frac'], l['roofline']['share_of_step'])"
timeout 300 python bench.py --workload c4 > $out/bench_c4_r2m.json 2> $out/bench_c4_r2m.err; python -c "
import json; l=json.loads(open('$out/bench_c4_r2m.json').read().strip().splitlines()[-1]); print('c4', l['value'], l['ms_per_pair'])"
python - <<PY
import time, numpy as np, sys
sys.path.insert(0, '.')
import lidarslam_ros2_b200 as m
from lidarslam_ros2_b200 import synth
src, tgt, _ = synth.registration_pair("headline", 2.0)
g = m.NormalDistributionsTransform(); g.setResolution(2.0)
for rep in range(3):
    t0 = time.perf_counter(); g.setInputTarget(tgt); t1 = time.perf_counter()
    print("setInputTarget pageable 1M: %.3f ms wall, device build %.3f ms" % (1e3*(t1-t0), g.stats()["target_build_ms"]))
import torch
pt = torch.from_numpy(tgt).pin_memory().numpy()
for rep in range(2):
    t0 = time.perf_counter(); g.setInputTarget(pt); t1 = time.perf_counter()
    print("setInputTarget pinned 1M: %.3f ms wall, device build %.3f ms" % (1e3*(t1-t0), g.stats()["target_build_ms"]))
PY
