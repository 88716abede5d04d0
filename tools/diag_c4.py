"""Developer diagnostic: where a loop-closure pair (config C4) spends its time, call by call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lidarslam_ros2_b200 as m
from lidarslam_ros2_b200 import synth

pairs = [next(iter(synth.loop_closure_pairs(8, first=i, count=1))) for i in range(4)]
ndt = m.NormalDistributionsTransform()
ndt.setResolution(2.0); ndt.setTransformationEpsilon(0.01); ndt.setMaximumIterations(100)
acc = {}
for rep in range(3):
    for _, src, tgt, _T in pairs:
        for name, fn in (("setInputTarget", lambda: ndt.setInputTarget(tgt)), ("setInputSource", lambda: ndt.setInputSource(src)),
                         ("align", lambda: ndt.align()), ("getFitnessScore", lambda: ndt.getFitnessScore())):
            t0 = time.perf_counter(); fn(); dt = time.perf_counter() - t0
            if rep: acc.setdefault(name, []).append(dt * 1e3)
print({k: round(float(np.median(v)), 3) for k, v in acc.items()}, "ms per call (median); n_src", len(pairs[0][1]), "n_tgt", len(pairs[0][2]))
print(ndt.stats())
