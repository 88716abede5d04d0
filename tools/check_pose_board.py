"""Under torchrun (one rank per GPU): every rank registers ITS scans against the same map with a pose board attached; the
poses every rank sees afterwards (stored by the peers' solver kernels over NVLink) must be bitwise what a plain
ncclAllGather of the ranks' batch results gives. Unequal counts per rank, several calls (tag parity), host and device
sources. Prints 'pose board ok' on rank 0."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import lidarslam_ros2_b200 as m
from lidarslam_ros2_b200 import batch, synth

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
src, tgt, _ = synth.registration_pair("small", 2.0)
g = m.NormalDistributionsTransform(device=local)
g.setResolution(2.0)
g.setTransformationEpsilon(0.01)
g.setNeighborhoodSearchMethod(m.DIRECT7)
g.setInputTarget(tgt)
comm = batch.RowComm(rank, world, local)
board = comm.create_board(32)
g.attachPoseBoard(board)
MAXN = 9
for call in range(6):
    n = 3 + (rank + call) % 5  # unequal counts
    rng = np.random.default_rng(1000 * rank + call)
    scans = []
    for k in range(n):
        s = src[rng.random(len(src)) < 0.9][:, :3].copy()
        s += rng.normal(0, 0.004, size=s.shape).astype(np.float32)
        scans.append(np.ascontiguousarray(s))
    if call % 2 == 0:
        r = g.alignBatch(scans)
    else:
        dev = [torch.from_numpy(np.concatenate([s, np.zeros((len(s), 1), np.float32)], axis=1)).cuda() for s in scans]
        r = g.alignBatchDevice([d.data_ptr() for d in dev], [len(d) for d in dev])
    poses, counts = g.gatheredPoses()
    # the same exchange through NCCL
    pad = np.zeros((MAXN, 17), dtype=np.float32)
    pad[:n, :16] = r["pose"].reshape(n, 16)
    pad[:n, 16] = 1.0
    ref = comm.all_gather_rows(pad).reshape(world, MAXN, 17)
    for q in range(world):
        nq = int(ref[q, :, 16].sum())
        assert counts[q] == nq, (call, q, counts, nq)
        assert np.array_equal(poses[q, :nq].reshape(nq, 16), ref[q, :nq, :16]), (call, q)
    assert np.array_equal(poses[rank, :n], r["pose"])
# latency of an attached call against a plain one (same scans): the exchange should hide inside the launch
g.attachPoseBoard(None)
t = []
for attach in (False, True, False, True):
    g.attachPoseBoard(board if attach else None)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    g.alignBatch(scans)
    t.append(1e3 * (time.perf_counter() - t0))
if rank == 0:
    print("batch call ms: plain %.3f, board %.3f, plain %.3f, board %.3f" % tuple(t))
    print("pose board ok")
g.attachPoseBoard(None)
board.close()
dist.destroy_process_group()
