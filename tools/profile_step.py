"""Short headline-workload run for ncu: a few aligns + a batched launch + fitness + voxel build (set target) so that
every kernel of the hot path appears in the launch list.  usage: profile_step.py [config] [n_align] [n_batch]"""
# (under ncu the host-buffer batch uploads everything before its launch: capi.cu, stream_write_value32)
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lidarslam_ros2_b200 as m
from lidarslam_ros2_b200 import synth
cfg = sys.argv[1] if len(sys.argv) > 1 else "headline"
n_align = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n_batch = int(sys.argv[3]) if len(sys.argv) > 3 else 8
src, tgt, _ = synth.registration_pair(cfg, 2.0)
g = m.NormalDistributionsTransform(); g.setResolution(2.0); g.setTransformationEpsilon(0.01)
g.setInputTarget(tgt); g.setInputSource(src)
for _ in range(n_align):
    T = g.align()
print("fitness", g.getFitnessScore(), "stats", g.stats())
if n_batch:
    import torch
    rng = np.random.default_rng(1)
    scans = [(src + rng.normal(0, 0.003, size=src.shape)).astype(np.float32) for _ in range(n_batch)]
    dev = [torch.from_numpy(np.concatenate([x, np.ones((len(x), 1), np.float32)], axis=1)).cuda() for x in scans]
    torch.cuda.synchronize()
    for _ in range(2):  # HBM-resident scans: the launch bench.py's `value` leg times
        r = g.alignBatchDevice([d.data_ptr() for d in dev], [d.shape[0] for d in dev])
    print("batch", r["iterations"], g.stats())
    r = g.alignBatch(scans)  # and once from host buffers (upload + the same kernel)
    print("batch from host", r["iterations"])
ds = m.voxel_grid_filter(src, 0.5)
print("voxelgrid", ds.shape)
