"""Where a loop-closure pair's time goes (BASELINE config 4: 32-ring scan vs 200k-pt submap, NDT res 2.0, max_iter 100):
wall clock per public call for pageable and pinned host buffers, and the two-engine sweep's per-pair time for both."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import lidarslam_ros2_b200 as m
from lidarslam_ros2_b200 import batch, synth

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
pairs = [next(iter(synth.loop_closure_pairs(64, first=i, count=1))) for i in range(n_pairs)]
srcs = [np.ascontiguousarray(p[1]) for p in pairs]
tgts = [np.ascontiguousarray(p[2]) for p in pairs]
psrcs = [torch.from_numpy(a).pin_memory().numpy() for a in srcs]
ptgts = [torch.from_numpy(a).pin_memory().numpy() for a in tgts]
g = m.NormalDistributionsTransform(); g.setResolution(2.0); g.setTransformationEpsilon(0.01); g.setMaximumIterations(100)
for name, S, T in (("pageable", srcs, tgts), ("pinned", psrcs, ptgts)):
    for rep in range(2):
        acc = np.zeros(4)
        for s, t in zip(S, T):
            t0 = time.perf_counter(); g.setInputTarget(t)
            t1 = time.perf_counter(); g.setInputSource(s)
            t2 = time.perf_counter(); g.align()
            t3 = time.perf_counter(); g.getFitnessScore()
            t4 = time.perf_counter()
            acc += [t1 - t0, t2 - t1, t3 - t2, t4 - t3]
    st = g.stats()
    print(f"{name:9s} per pair [ms]: setInputTarget {1e3*acc[0]/len(S):.3f} (device build {st['target_build_ms']:.3f})  setInputSource {1e3*acc[1]/len(S):.3f}  "
          f"align {1e3*acc[2]/len(S):.3f} (kernel {st['solve_ms']:.3f})  getFitnessScore {1e3*acc[3]/len(S):.3f}  total {1e3*acc.sum()/len(S):.3f}", flush=True)
sw = batch.LoopSweep(m, device=0, resolution=2.0, max_iterations=100)
for name, S, T in (("pageable", srcs, tgts), ("pinned", psrcs, ptgts)):
    sw.run(S, T, list(range(len(S))))
    t0 = time.perf_counter()
    for _ in range(3):
        sw.run(S, T, list(range(len(S))))
    print(f"sweep {name:9s}: {1e3*(time.perf_counter()-t0)/(3*len(S)):.3f} ms per pair", flush=True)
