"""Where the cycles of a batched launch go (B200REG_BATCH_PROFILE=1): per evaluator CTA the SM cycles spent waiting for control
blocks / evaluating / reducing, per controller CTA the cycles waiting for rows / in the controller step."""
import os, sys
os.environ["B200REG_BATCH_PROFILE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import numpy as np
import torch
import lidarslam_ros2_b200 as m
from lidarslam_ros2_b200 import _capi, synth
K = int(sys.argv[1]) if len(sys.argv) > 1 else 20
src, tgt, _ = synth.registration_pair("headline", 2.0)
rng = np.random.default_rng(1)
scans = [(src + rng.normal(0, 0.003, size=src.shape)).astype(np.float32) for _ in range(K)]
g = m.NormalDistributionsTransform(); g.setResolution(2.0); g.setTransformationEpsilon(0.01); g.setInputTarget(tgt)
L = _capi.lib()
L.b200reg_debug_cta_eval_ns.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
dev = [torch.from_numpy(np.concatenate([x, np.ones((len(x), 1), np.float32)], axis=1)).cuda() for x in scans]
torch.cuda.synchronize()
ptrs, cnts = [d.data_ptr() for d in dev], [d.shape[0] for d in dev]
for slots in (3, 2, 1):
    g.setBatchSlots(slots)
    g.alignBatchDevice(ptrs, cnts)
    r = g.alignBatchDevice(ptrs, cnts)
    st = g.stats()
    buf = np.zeros((256, 4), dtype=np.uint32)
    L.b200reg_debug_cta_eval_ns(g._h, buf.ctypes.data, 256)
    n_eval = st["grid_ctas"] - slots
    ev = buf[:n_eval].astype(np.float64) * 1024 / 1.965e3  # -> microseconds at 1.965 GHz
    ct = buf[n_eval:n_eval + slots].astype(np.float64) * 1024 / 1.965e3
    evals = int(r["evaluations"].sum())
    print(f"slots {slots}: kernel {st['solve_ms']*1e3:.0f} us, {evals} evaluations -> {st['solve_ms']*1e3/evals:.2f} us/evaluation")
    print(f"  evaluator CTAs (mean / max over {n_eval}) [us]: wait {ev[:,0].mean():.0f}/{ev[:,0].max():.0f}  evaluate {ev[:,1].mean():.0f}/{ev[:,1].max():.0f}  "
          f"reduce {ev[:,2].mean():.0f}/{ev[:,2].max():.0f}  -> per evaluation: wait {ev[:,0].mean()/evals:.2f} evaluate {ev[:,1].mean()/evals:.2f} reduce {ev[:,2].mean()/evals:.2f}")
    for s in range(slots):
        print(f"  controller {s} [us]: waiting for rows {ct[s,0]:.0f}  step {ct[s,1]:.0f}")
