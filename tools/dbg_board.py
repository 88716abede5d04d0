import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import lidarslam_ros2_b200 as m
from lidarslam_ros2_b200 import batch, synth
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_pose_board import _problem
g, scans = _problem(m, 7)
plain = g.alignBatch(scans)
comm = batch.RowComm(0, 1, 0)
board = comm.create_board(16)
g.attachPoseBoard(board)
dev = [torch.from_numpy(np.concatenate([s, np.zeros((len(s), 1), np.float32)], axis=1)).cuda() for s in scans]
for mode in ("device", "host", "host", "device"):
    t0 = time.time()
    try:
        if mode == "device":
            r = g.alignBatchDevice([d.data_ptr() for d in dev], [len(d) for d in dev])
        else:
            r = g.alignBatch(scans)
        p, c = g.gatheredPoses()
        print(mode, "ok", c, np.array_equal(p[0], r["pose"]), "%.3f s" % (time.time() - t0))
    except Exception as e:
        print(mode, "FAILED", e, "%.3f s" % (time.time() - t0))
