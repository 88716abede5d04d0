"""Short GICP run (BASELINE config 3 shape) for the ncu launch list and a host-side wall-clock split."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lidarslam_ros2_b200 as m
from lidarslam_ros2_b200 import synth
n_align = int(sys.argv[1]) if len(sys.argv) > 1 else 2
src, tgt, _ = synth.registration_pair("headline", 2.0)
g = m.GeneralizedIterativeClosestPoint(); g.setMaxCorrespondenceDistance(5.0)
t0 = time.perf_counter(); g.setInputTarget(tgt); g.setInputSource(src); t1 = time.perf_counter()
T = g.align(); t2 = time.perf_counter()
print("set clouds %.2f ms, first align (target covariances) %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
for _ in range(n_align):
    t0 = time.perf_counter(); g.setInputSource(src); t1 = time.perf_counter(); T = g.align(); t2 = time.perf_counter()
    print("setInputSource %.2f ms, align %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), "stats", g.stats())
