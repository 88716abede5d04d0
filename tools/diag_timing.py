"""Developer diagnostic: per-phase timing inside the persistent NDT solver (B200REG_TIMING=1) and entrywise
parity of (g, H) against the oracle. Not part of the product path."""
import ctypes as C
import os
import sys

os.environ["B200REG_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import lidarslam_ros2_b200 as m
import oracle
from lidarslam_ros2_b200 import _capi, synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
res = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
src, tgt, _ = synth.registration_pair(cfg, res)
g = m.NormalDistributionsTransform()
g.setResolution(res)
g.setTransformationEpsilon(0.01)
g.setInputTarget(tgt)
g.setInputSource(src)
for _ in range(int(os.environ.get("DIAG_WARM", "3"))):
    T = g.align()
st = g.stats()
print("stats", st)
buf = np.zeros((48, 10), dtype=np.uint64)
L = _capi.lib()
L.b200reg_debug_timing.argtypes = [C.c_void_p, C.c_void_p]
L.b200reg_debug_timing(g._h, buf.ctypes.data_as(C.c_void_p))
E = st["evaluations"]
t = buf[:E].astype(np.int64)
names = ["start", "eval_done", "partial_written", "arrived", "last_detected", "partials_reduced", "controller_done", "released"]
print("per-round phase durations [us] (CTA0 unless noted):")
rows = []
for r in range(E):
    x = t[r]
    rows.append([(x[1] - x[0]), (x[2] - x[1]), (x[3] - x[2]), (x[4] - x[3]), (x[5] - x[4]), (x[6] - x[5]), (x[7] - x[6]),
                 (t[r + 1][0] - x[7]) if r + 1 < E else 0, (x[7] - x[0])])
rows = np.array(rows) / 1e3
hdr = ["evaluate", "cta_reduce", "fence+arrive", "wait_last(all arrive)", "reduce_partials", "controller", "release->seen",
       "ctl_read->next", "round_total"]
print(" ".join(f"{h:>22s}" for h in hdr))
for r in rows:
    print(" ".join(f"{v:22.2f}" for v in r))
print("median:")
print(" ".join(f"{v:22.2f}" for v in np.median(rows[1:-1], axis=0)))
dry1 = (t[:, 8] - t[:, 5]) / 1e3
dry2 = (t[:, 9] - t[:, 8]) / 1e3
real = (t[:, 6] - t[:, 9]) / 1e3
print("controller_fast dry pass 1 us:", np.round(dry1[1:-1], 2))
print("controller_fast dry pass 2 us:", np.round(dry2[1:-1], 2))
print("real pass + build_control + publish us:", np.round(real[1:-1], 2))

nc = st["grid_ctas"] - 1
ce = np.zeros(nc, dtype=np.uint32)
L.b200reg_debug_cta_eval_ns.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
L.b200reg_debug_cta_eval_ns(g._h, ce.ctypes.data_as(C.c_void_p), nc)
ce = ce / 1e3
print("per-CTA evaluate us (round 2): min %.2f p10 %.2f median %.2f p90 %.2f max %.2f" % (ce.min(), np.percentile(ce, 10), np.median(ce), np.percentile(ce, 90), ce.max()))
print("  by CTA index (every 16th):", np.round(ce[::16], 2))
print("  slowest CTAs:", np.argsort(ce)[-8:], np.round(np.sort(ce)[-8:], 2))

# entrywise (g, H) parity with per-entry scale sqrt(|Hii Hjj|)
o = oracle.NDT(resolution=res, transformation_epsilon=0.01)
o.set_target(tgt)
o.set_source(src)
To = o.align()
print("pose err gpu vs cpu:", synth.pose_error(T, To), "iters", g.getFinalNumIteration(), o.iterations)
for p in (np.zeros(6), np.array([0.38, -0.24, 0.06, 0.007, -0.005, 0.026])):
    Tm = oracle.pose_to_matrix(p)
    sg, gg, Hg = g.derivatives(Tm, p)
    so, go, Ho = o.derivatives(Tm, p)
    d = np.sqrt(np.abs(np.diag(Ho)))
    rel = np.abs(Hg - Ho) / np.outer(d, d)
    print("p", p, "score rel", abs(sg - so) / abs(so), "max |dH|/sqrt(HiiHjj)", rel.max(), "max |dg|/|g|", np.abs(gg - go).max() / np.abs(go).max())
    print("  g gpu", gg)
    print("  g cpu", go)
    w, _ = np.linalg.eigh(Ho)
    print("  eig(H cpu)", w)
