"""Developer diagnostic: per-phase timing inside the persistent NDT solver (B200REG_TIMING=1) and entrywise
parity of (g, H) against the oracle. Not part of the product path."""
import ctypes as C
import os
import sys

os.environ["B200REG_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
buf = np.zeros((48, 12), dtype=np.uint64)
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
print("reducing warps finished (us before warp 0 noticed): first", np.round((t[1:-1, 4] - t[1:-1, 10]) / 1e3, 2), "last", np.round((t[1:-1, 4] - t[1:-1, 11]) / 1e3, 2))
print("CTA0 published -> last reducing warp done (us):", np.round((t[1:-1, 11].astype(np.int64) - t[1:-1, 2].astype(np.int64)) / 1e3, 2))
print("controller: step", np.round((t[1:-1, 8] - t[1:-1, 5]) / 1e3, 2), "build_control", np.round((t[1:-1, 9] - t[1:-1, 8]) / 1e3, 2), "publish", np.round((t[1:-1, 6] - t[1:-1, 9]) / 1e3, 2))

print("kernel prologue (entry -> first round) us:", (int(buf[47][1]) - int(buf[47][0])) / 1e3,
      " entry -> last publish us:", (int(t[E - 1][6]) - int(buf[47][0])) / 1e3, " solve_ms(events) us:", st["solve_ms"] * 1e3)
import time
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(200):
        g.align()
    dt = (time.perf_counter() - t0) / 200
    print("host wall per align() us: %.1f   solve_ms us: %.1f" % (dt * 1e6, g.stats()["solve_ms"] * 1e3))

nc = st["grid_ctas"] - 1
ce = np.zeros((nc, 4), dtype=np.uint32)
L.b200reg_debug_cta_eval_ns.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
L.b200reg_debug_cta_eval_ns(g._h, ce.ctypes.data_as(C.c_void_p), nc)
rel = np.uint32(int(t[1][6]) & 0xffffffff)   # controller published round 1's control block -> round 2 starts
det = ((np.uint32(int(t[2][4]) & 0xffffffff) - rel).astype(np.int32)) / 1e3
start = (ce[:, 0] - rel).astype(np.int32) / 1e3
evend = (ce[:, 1] - rel).astype(np.int32) / 1e3
pub = (ce[:, 2] - rel).astype(np.int32) / 1e3
def q(name, v):
    print(f"  {name:28s} min {v.min():6.2f} p10 {np.percentile(v,10):6.2f} med {np.median(v):6.2f} p90 {np.percentile(v,90):6.2f} max {v.max():6.2f}  (argmax {int(np.argmax(v))})")
print(f"round 2, per evaluator CTA, us after the controller's release (controller saw all rows at {det:.2f}):")
q("start (gen seen)", start); q("evaluate end", evend); q("published", pub)
q("evaluate duration", evend - start); q("cta reduce+publish", pub - evend)
late = np.argsort(pub)[-6:]
print("  latest publishers:", late, "start", np.round(start[late], 2), "eval", np.round((evend - start)[late], 2), "pub", np.round(pub[late], 2))

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
