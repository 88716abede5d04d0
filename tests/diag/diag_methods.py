import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import lidarslam_ros2_b200 as m
import oracle
from lidarslam_ros2_b200 import synth
src, tgt, _ = synth.registration_pair("small", 2.0)
for method in (2, 3, 1, 0):
    g = m.NormalDistributionsTransform(); g.setResolution(2.0); g.setNeighborhoodSearchMethod(method)
    g.setInputTarget(tgt); g.setInputSource(src)
    o = oracle.NDT(resolution=2.0, search_method=method); o.set_target(tgt); o.set_source(src)
    for p in (np.zeros(6), np.array([0.21, -0.13, 0.04, 0.006, -0.004, 0.02]), np.array([-0.4, 0.3, -0.1, 2.9, 0.01, -0.3])):
        T = oracle.pose_to_matrix(p)
        for hess in (True, False):
            sg, gg, Hg = g.derivatives(T, p, hess); hits = g.stats()["hits"]
            so, go, Ho = o.derivatives(T, p, hess)
            print(method, hess, "score", sg, so, "d", sg - so, "hits", hits, "dg", np.abs(gg - go).max() / max(1, np.abs(go).max()))
