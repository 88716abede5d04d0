"""GICP parity diagnosis on the GPU box: where do the GPU path and the CPU oracle part ways?
 usage: diag_gicp.py [tiny|small|c1|c2 ...]
For each config: covariance agreement (source / target) with the eigen-gap of the mismatching points, then the pose after
1, 2, 3, 5, 10, 30, 100 outer iterations (transformation_epsilon 1e-8 like the node, sm.cpp:119) on both sides."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import lidarslam_ros2_b200 as m
import oracle
from lidarslam_ros2_b200 import synth

oracle.build()
for cfg in (sys.argv[1:] or ["tiny", "small", "c1"]):
    src, tgt, T_gt = synth.registration_pair(cfg, 2.0)
    print(f"== {cfg}: {len(src)} vs {len(tgt)}", flush=True)
    g = m.GeneralizedIterativeClosestPoint(); g.setMaxCorrespondenceDistance(5.0); g.setTransformationEpsilon(1e-8)
    o = oracle.GICP(max_correspondence_distance=5.0, transformation_epsilon=1e-8)
    g.setInputTarget(tgt); g.setInputSource(src); o.set_target(tgt); o.set_source(src)
    for its in (1, 2, 3, 5, 10, 30, 100):
        g.setMaximumIterations(its); o.set("max_iterations", its)
        t0 = time.time(); Tg = g.align(); tg = time.time() - t0
        t0 = time.time(); To = o.align(); to = time.time() - t0
        dt, dr = synth.pose_error(Tg, To)
        et = synth.pose_error(Tg, T_gt)
        print(f"  max_it {its:3d}: gpu it={g.stats()['iterations']} conv={g.hasConverged()} corr={g.numCorrespondences()} | "
              f"cpu it={o.iterations} conv={o.converged} corr={o.num_correspondences()} | dT={dt:.2e} m {dr:.2e} rad | "
              f"gpu-vs-truth {et[0]:.3e} {et[1]:.3e} | {tg*1e3:.1f} ms / {to*1e3:.0f} ms", flush=True)
        if its == 1:
            for which in ("source", "target"):
                cg, co = g.covariances(which), o.covariances(which)
                err = np.abs(cg - co).max(axis=(1, 2))
                bad = np.flatnonzero(err > 1e-6)
                print(f"  cov {which}: n={len(err)} mismatching(>1e-6)={len(bad)} ({100*len(bad)/len(err):.3f} %) max={err.max():.2e}")
    # default epsilon (gicp_omp.h:118: 5e-4) as the oracle test uses
    g2 = m.GeneralizedIterativeClosestPoint(); g2.setMaxCorrespondenceDistance(5.0)
    o2 = oracle.GICP(max_correspondence_distance=5.0)
    g2.setInputTarget(tgt); g2.setInputSource(src); o2.set_target(tgt); o2.set_source(src)
    Tg, To = g2.align(), o2.align()
    print(f"  eps 5e-4 default: gpu it={g2.stats()['iterations']} cpu it={o2.iterations} dT={synth.pose_error(Tg, To)}", flush=True)
