"""Latency of b200comm_all_gather_rows (ncclAllGather from C) per call, under torchrun: first calls vs steady state."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from lidarslam_ros2_b200 import batch
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = batch.RowComm(rank, world, local)
rows = np.full((20, 16), float(rank), dtype=np.float32)
ts = []
for k in range(60):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter(); out = comm.all_gather_rows(rows); ts.append(1e6 * (time.perf_counter() - t0))
assert out.shape == (world * 20, 16) and out[-1, 0] == world - 1
if rank == 0:
    print("b200comm all_gather_rows us: first 5", [round(t) for t in ts[:5]], "median of the rest", round(float(np.median(ts[5:]))), "max", round(max(ts[5:])))
# the same call after the links / the GPU sat idle for a few milliseconds (bench.py's timed region: one 1.4 ms kernel first)
for gap_ms in (0.5, 2.0, 5.0):
    ts = []
    for k in range(30):
        dist.barrier(); torch.cuda.synchronize()
        time.sleep(gap_ms * 1e-3)
        t0 = time.perf_counter(); out = comm.all_gather_rows(rows); ts.append(1e6 * (time.perf_counter() - t0))
    if rank == 0:
        print("  after %.1f ms of idle: median %d us, min %d, max %d" % (gap_ms, round(float(np.median(ts))), round(min(ts)), round(max(ts))))
# ... and after a busy GPU (a spinning kernel of ~1.5 ms on the current stream) instead of an idle one
x = torch.zeros(1 << 20, device="cuda")
ts = []
for k in range(30):
    dist.barrier(); torch.cuda.synchronize()
    torch.cuda._sleep(3_000_000)  # ~1.5 ms of device spin
    torch.cuda.synchronize()
    t0 = time.perf_counter(); out = comm.all_gather_rows(rows); ts.append(1e6 * (time.perf_counter() - t0))
if rank == 0:
    print("  after a 1.5 ms spinning kernel: median %d us, min %d, max %d" % (round(float(np.median(ts))), round(min(ts)), round(max(ts))))
g = torch.zeros((world * 20, 16), device="cuda"); x = torch.zeros((20, 16), device="cuda")
ts = []
for k in range(60):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter(); dist.all_gather_into_tensor(g, x); torch.cuda.synchronize(); ts.append(1e6 * (time.perf_counter() - t0))
if rank == 0:
    print("torch all_gather_into_tensor us: first 5", [round(t) for t in ts[:5]], "median of the rest", round(float(np.median(ts[5:]))))
dist.destroy_process_group()
