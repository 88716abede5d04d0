"""Two loop-closure pairs (BASELINE config 4) through the public call sequence, for an ncu launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lidarslam_ros2_b200 as m
from lidarslam_ros2_b200 import synth
pairs = [next(iter(synth.loop_closure_pairs(64, first=i, count=1))) for i in (5, 40)]
g = m.NormalDistributionsTransform(); g.setResolution(2.0); g.setTransformationEpsilon(0.01); g.setMaximumIterations(100)
for rep in range(2):
    for _, src, tgt, _T in pairs:
        g.setInputTarget(tgt); g.setInputSource(src); g.align()
        print("fitness", g.getFitnessScore(), g.stats()["iterations"])
