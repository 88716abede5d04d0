import os, sys, time, subprocess, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lidarslam_ros2_b200 as m
from lidarslam_ros2_b200 import synth
cfg, res = sys.argv[1], float(sys.argv[2])
src, tgt, _ = synth.registration_pair(cfg, res)
g = m.NormalDistributionsTransform(); g.setResolution(res); g.setTransformationEpsilon(0.01)
g.setInputTarget(tgt); g.setInputSource(src)
rows = []
def smi():
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
    for line in p.stdout:
        rows.append((time.perf_counter(), line.strip()))
        if stop[0]:
            break
    p.terminate()
stop = [False]
t = threading.Thread(target=smi, daemon=True); t.start()
time.sleep(0.5)
t0 = time.perf_counter()
ms = []
for k in range(6000):
    g.align()
    ms.append((time.perf_counter() - t0, g.stats()["solve_ms"], g.stats()["evaluations"]))
stop[0] = True
time.sleep(0.2)
ms = np.array(ms)
for a, b in ((0, 10), (10, 100), (100, 500), (500, 1500), (1500, 3000), (3000, 6000)):
    seg = ms[a:b]
    print(f"aligns {a}-{b}: t={seg[0,0]:.3f}..{seg[-1,0]:.3f}s solve_ms median {np.median(seg[:,1]):.4f} us/eval {1e3*np.median(seg[:,1]/seg[:,2]):.2f} wall/align {1e3*(seg[-1,0]-seg[0,0])/len(seg):.4f} ms")
print("smi samples (t rel, sm clk, max, W, reasons):")
for tt, l in rows[::max(1, len(rows)//25)]:
    print(f"  {tt - t0:7.3f}  {l}")
