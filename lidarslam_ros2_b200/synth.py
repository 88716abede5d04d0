"""Deterministic synthetic "street canyon" clouds shared by tests, bench and oracle.

The reference ships no benchmark data beyond two PCD scans (Thirdparty/ndt_omp_ros2/data/*.pcd, used by
apps/align.cpp:42-127); BASELINE.json's configs are defined on synthetic clouds of named sizes. This module
generates them with a counter-based RNG (splitmix64, base seed 0x20260922) so that the CPU oracle and the
CUDA engine read bit-identical float32 buffers on every machine.

Frames: the *map* (registration target) is expressed in the nominal sensor frame S0 = Trans(0, 0, 1.9); a
*scan* (registration source) is ray-cast from the true sensor pose S0 * T_gt and expressed in its own sensor
frame, so that  p_map = T_gt * p_scan  and align(guess = I) should return T_gt.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

BASE_SEED = 0x20260922
SENSOR_HEIGHT = 1.9
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


class Rng:
    """Counter-based stream: value k of stream s depends only on (BASE_SEED, s, k)."""

    def __init__(self, stream: int):
        s = _splitmix64(np.array([BASE_SEED], dtype=np.uint64))
        self._key = _splitmix64(s ^ _splitmix64(np.array([stream], dtype=np.uint64)))[0]
        self._ctr = 0

    def uniform(self, n: int) -> np.ndarray:
        idx = np.arange(self._ctr, self._ctr + n, dtype=np.uint64)
        self._ctr += n
        with np.errstate(over="ignore"):
            bits = _splitmix64((idx * np.uint64(0xD1342543DE82EF95) + self._key) & _M64)
        return (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)

    def normal(self, n: int) -> np.ndarray:
        u1 = np.maximum(self.uniform(n), 1e-300)
        u2 = self.uniform(n)
        return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * math.pi * u2)


def rpy_matrix(rx: float, ry: float, rz: float) -> np.ndarray:
    """R = Rx(rx) * Ry(ry) * Rz(rz) — the NDT parameterisation (ndt_omp_impl.hpp:146-149)."""
    cx, sx, cy, sy, cz, sz = math.cos(rx), math.sin(rx), math.cos(ry), math.sin(ry), math.cos(rz), math.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rx @ Ry @ Rz


def pose_matrix(t, rpy) -> np.ndarray:
    T = np.eye(4)
    T[:3, :3] = rpy_matrix(*rpy)
    T[:3, 3] = t
    return T


def ground_truth(resolution: float) -> np.ndarray:
    """T_gt of SURVEY.md §8d: t=(0.40,-0.25,0.06) m, rpy=(0.4,-0.3,1.5) deg for res >= 2; halved for res 1."""
    s = 1.0 if resolution >= 2.0 else 0.5
    d = math.pi / 180.0
    return pose_matrix((0.40 * s, -0.25 * s, 0.06 * s), (0.4 * d * s, -0.3 * d * s, 1.5 * d * s))


@dataclass
class Scene:
    boxes: np.ndarray      # (B, 6): xmin, ymin, zmin, xmax, ymax, zmax  (world frame, ground z = 0)
    cylinders: np.ndarray  # (C, 4): cx, cy, radius, height
    x_range: tuple = (-150.0, 150.0)
    y_range: tuple = (-40.0, 40.0)
    facade_y: float = 18.0
    facade_h: float = 14.0


def make_scene(stream: int = 1) -> Scene:
    r = Rng(stream)
    boxes = []
    # pilasters: 1.5 m deep, 1.0 m wide, every 7 m on both facades
    for k in range(-21, 22):
        xc = 7.0 * k
        boxes.append([xc - 0.5, 18.0 - 1.5, 0.0, xc + 0.5, 18.0, 14.0])
        boxes.append([xc - 0.5 + 3.5, -18.0, 0.0, xc + 0.5 + 3.5, -18.0 + 1.5, 14.0])
    u = r.uniform(40 * 6).reshape(40, 6)
    for k in range(40):
        lx = 1.8 + 2.7 * u[k, 0]
        ly = 1.6 + 0.6 * u[k, 1]
        lz = 1.4 + 1.6 * u[k, 2]
        cx = -140.0 + 280.0 * u[k, 3]
        cy = -14.0 + 28.0 * u[k, 4]
        if abs(cx) < 4.0 and abs(cy) < 3.0:  # keep the sensor's own spot free
            cy += 7.0
        boxes.append([cx - lx / 2, cy - ly / 2, 0.0, cx + lx / 2, cy + ly / 2, lz])
    u = r.uniform(60 * 4).reshape(60, 4)
    cyl = []
    for k in range(60):
        cx = -140.0 + 280.0 * u[k, 0]
        cy = (-15.5 if u[k, 1] < 0.5 else 15.5) + (u[k, 1] - 0.5) * 2.0
        if abs(cx) < 3.0 and abs(cy) < 3.0:
            cx += 6.0
        cyl.append([cx, cy, 0.1 + 0.3 * u[k, 2], 3.0 + 6.0 * u[k, 3]])
    return Scene(boxes=np.array(boxes), cylinders=np.array(cyl))


# --------------------------------------------------------------------------------------------------
# map sampling (area-uniform over all surfaces)
# --------------------------------------------------------------------------------------------------
def sample_map(scene: Scene, n: int, stream: int, x_window: tuple | None = None, noise: float = 0.01) -> np.ndarray:
    """n points area-uniform on the scene surfaces (world frame shifted to the nominal sensor frame S0).

    x_window=(x0, x1) restricts the sample to a stretch of the canyon (local maps of config 4).
    Returns float32 (n, 3).
    """
    r = Rng(stream)
    x0, x1 = x_window if x_window is not None else scene.x_range
    y0, y1 = scene.y_range
    bx = scene.boxes[(scene.boxes[:, 3] > x0) & (scene.boxes[:, 0] < x1)]
    cy = scene.cylinders[(scene.cylinders[:, 0] > x0) & (scene.cylinders[:, 0] < x1)]
    surf_area = [(x1 - x0) * (y1 - y0), (x1 - x0) * scene.facade_h, (x1 - x0) * scene.facade_h]
    lx, ly, lz = bx[:, 3] - bx[:, 0], bx[:, 4] - bx[:, 1], bx[:, 5] - bx[:, 2]
    box_faces = np.stack([lx * ly, lx * lz, lx * lz, ly * lz, ly * lz], axis=1)  # top, y-, y+, x-, x+
    cyl_area = 2 * math.pi * cy[:, 2] * cy[:, 3]
    areas = np.concatenate([np.array(surf_area), box_faces.reshape(-1), cyl_area])
    cdf = np.cumsum(areas) / areas.sum()
    pick = np.searchsorted(cdf, r.uniform(n), side="right").clip(0, len(areas) - 1)
    a, b = r.uniform(n), r.uniform(n)
    pts = np.zeros((n, 3))
    m = pick == 0  # ground
    pts[m] = np.stack([x0 + (x1 - x0) * a[m], y0 + (y1 - y0) * b[m], np.zeros(m.sum())], axis=1)
    for s, ysign in ((1, 1.0), (2, -1.0)):
        m = pick == s
        pts[m] = np.stack([x0 + (x1 - x0) * a[m], np.full(m.sum(), ysign * scene.facade_y), scene.facade_h * b[m]], axis=1)
    nb = len(bx)
    m = (pick >= 3) & (pick < 3 + 5 * nb)
    if m.any():
        bi = (pick[m] - 3) // 5
        fi = (pick[m] - 3) % 5
        B = bx[bi]
        aa, bb = a[m], b[m]
        x = B[:, 0] + (B[:, 3] - B[:, 0]) * aa
        y = B[:, 1] + (B[:, 4] - B[:, 1]) * bb
        z = np.where(fi == 0, B[:, 5], B[:, 2] + (B[:, 5] - B[:, 2]) * bb)
        y = np.where(fi == 1, B[:, 1], np.where(fi == 2, B[:, 4], y))
        yy = B[:, 1] + (B[:, 4] - B[:, 1]) * aa
        x = np.where(fi == 3, B[:, 0], np.where(fi == 4, B[:, 3], x))
        y = np.where(fi >= 3, yy, y)
        pts[m] = np.stack([x, y, z], axis=1)
    m = pick >= 3 + 5 * nb
    if m.any():
        ci = pick[m] - (3 + 5 * nb)
        Cc = cy[ci]
        th = 2 * math.pi * a[m]
        pts[m] = np.stack([Cc[:, 0] + Cc[:, 2] * np.cos(th), Cc[:, 1] + Cc[:, 2] * np.sin(th), Cc[:, 3] * b[m]], axis=1)
    pts[:, 2] += np.where(pick == 0, 0.02 * r.normal(n), 0.0)  # ground height noise sigma = 0.02
    pts += noise * np.stack([r.normal(n), r.normal(n), r.normal(n)], axis=1)
    pts[:, 2] -= SENSOR_HEIGHT  # express in the nominal sensor frame S0
    return pts.astype(np.float32)


# --------------------------------------------------------------------------------------------------
# scan ray casting
# --------------------------------------------------------------------------------------------------
def _ray_cast(scene: Scene, o: np.ndarray, d: np.ndarray, max_range: float) -> np.ndarray:
    """o: (3,) origin, d: (n,3) unit directions (world). Returns range (inf where nothing hit)."""
    n = len(d)
    best = np.full(n, np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        # ground z = 0
        t = -o[2] / d[:, 2]
        hx, hy = o[0] + t * d[:, 0], o[1] + t * d[:, 1]
        ok = (t > 0) & (hx >= scene.x_range[0]) & (hx <= scene.x_range[1]) & (hy >= scene.y_range[0]) & (hy <= scene.y_range[1])
        best = np.where(ok & (t < best), t, best)
        for ys in (scene.facade_y, -scene.facade_y):
            t = (ys - o[1]) / d[:, 1]
            hx, hz = o[0] + t * d[:, 0], o[2] + t * d[:, 2]
            ok = (t > 0) & (hx >= scene.x_range[0]) & (hx <= scene.x_range[1]) & (hz >= 0) & (hz <= scene.facade_h)
            best = np.where(ok & (t < best), t, best)
        inv = 1.0 / d
        chunk = 16384
        for s in range(0, n, chunk):
            e = min(n, s + chunk)
            iv = inv[s:e, None, :]
            t1 = (scene.boxes[None, :, 0:3] - o[None, None, :]) * iv
            t2 = (scene.boxes[None, :, 3:6] - o[None, None, :]) * iv
            tmin = np.minimum(t1, t2).max(axis=2)
            tmax = np.maximum(t1, t2).min(axis=2)
            hit = (tmax >= np.maximum(tmin, 0.0)) & (tmin > 0)
            tb = np.where(hit, tmin, np.inf).min(axis=1)
            best[s:e] = np.minimum(best[s:e], tb)
            # cylinders (vertical, finite height, side surface only)
            dx, dy = d[s:e, 0:1], d[s:e, 1:2]
            ox = o[0] - scene.cylinders[None, :, 0]
            oy = o[1] - scene.cylinders[None, :, 1]
            A = dx * dx + dy * dy
            Bq = 2 * (ox * dx + oy * dy)
            Cq = ox * ox + oy * oy - scene.cylinders[None, :, 2] ** 2
            disc = Bq * Bq - 4 * A * Cq
            tc = (-Bq - np.sqrt(np.where(disc > 0, disc, np.nan))) / (2 * A)
            hz = o[2] + tc * d[s:e, 2:3]
            okc = (disc > 0) & (tc > 0) & (hz >= 0) & (hz <= scene.cylinders[None, :, 3])
            tcm = np.where(okc, tc, np.inf).min(axis=1)
            best[s:e] = np.minimum(best[s:e], tcm)
    best[best > max_range] = np.inf
    return best


def make_scan(scene: Scene, rings: int, azimuths: int, sensor_pose_world: np.ndarray, stream: int,
              max_range: float = 100.0, range_noise: float = 0.02) -> np.ndarray:
    """Ray-cast a rings x azimuths LiDAR scan from sensor_pose_world (4x4). Points in the sensor frame, float32."""
    r = Rng(stream)
    el = np.deg2rad(np.linspace(-25.0, 15.0, rings))
    az = np.arange(azimuths) * (2 * math.pi / azimuths)
    E, A = np.meshgrid(el, az, indexing="ij")
    ds = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], axis=-1).reshape(-1, 3)
    R, t = sensor_pose_world[:3, :3], sensor_pose_world[:3, 3]
    dw = ds @ R.T
    rng = _ray_cast(scene, t, dw, max_range)
    rng = rng + range_noise * r.normal(len(rng))
    keep = np.isfinite(rng) & (rng > 0.5)
    return (ds[keep] * rng[keep, None]).astype(np.float32)


def sensor_pose(T_rel: np.ndarray, x_along: float = 0.0) -> np.ndarray:
    """World pose of a sensor whose pose relative to the nominal frame S0(x_along) is T_rel."""
    S0 = np.eye(4)
    S0[:3, 3] = (x_along, 0.0, SENSOR_HEIGHT)
    return S0 @ T_rel


def _drive_world_pose(k: int, step: float, x_start: float, yaw_amp_deg: float) -> np.ndarray:
    d = math.pi / 180.0
    yaw = yaw_amp_deg * d * math.sin(2.0 * math.pi * k / 40.0)
    T = pose_matrix((step * k, 0.3 * math.sin(2.0 * math.pi * k / 60.0), 0.0), (0.0, 0.0, yaw))
    return sensor_pose(T, x_start)


def drive_frame(k: int, rings: int = 16, azimuths: int = 625, step: float = 0.5, x_start: float = -60.0,
                yaw_amp_deg: float = 1.0, stream: int = 7000):
    """Frame k of the config-5 drive (reproducible on its own): (scan in the sensor frame, T_rel_k) where T_rel_k is the
    k-th sensor pose relative to the first one — the ground truth of the frontend's k-th pose, which starts at identity."""
    scene = make_scene()
    S0 = _drive_world_pose(0, step, x_start, yaw_amp_deg)
    Sk = _drive_world_pose(k, step, x_start, yaw_amp_deg)
    scan = make_scan(scene, rings, azimuths, Sk, stream=stream + k)
    return scan, np.linalg.inv(S0) @ Sk


def drive_stream(n_frames: int, rings: int = 16, azimuths: int = 625, step: float = 0.5, x_start: float = -60.0,
                 yaw_amp_deg: float = 1.0, stream: int = 7000, workers: int = 1):
    """Config 5: a sensor driving down the canyon, `step` metres per frame with a slow yaw weave and a small lateral
    drift. Yields (scan, T_rel_k) per frame; workers > 1 ray-casts the frames in a process pool."""
    if workers <= 1:
        for k in range(n_frames):
            yield drive_frame(k, rings, azimuths, step, x_start, yaw_amp_deg, stream)
        return
    from concurrent.futures import ProcessPoolExecutor

    with ProcessPoolExecutor(max_workers=workers) as ex:
        futs = [ex.submit(drive_frame, k, rings, azimuths, step, x_start, yaw_amp_deg, stream) for k in range(n_frames)]
        for f in futs:
            yield f.result()


_SCAN_SHAPES = {"10k": (16, 625), "60k": (32, 1875), "100k": (64, 1563)}


def registration_pair(config: str, resolution: float = 2.0):
    """(source, target, T_gt) for a named BASELINE configuration.

    config: "c1" 10k/50k, "c2" 60k/500k, "headline" 100k/1M, "tiny" 2k/20k (unit tests).
    """
    table = {
        "tiny": (8, 300, 20_000, 7),
        "small": (16, 313, 50_000, 6),
        "c1": (16, 625, 50_000, 1),
        "c2": (32, 1875, 500_000, 2),
        "headline": (64, 1563, 1_000_000, 3),
    }
    rings, azim, n_tgt, cid = table[config]
    scene = make_scene()
    T_gt = ground_truth(resolution)
    tgt = sample_map(scene, n_tgt, stream=cid * 1000 + 1)
    src = make_scan(scene, rings, azim, sensor_pose(T_gt), stream=cid * 1000 + 2)
    return src, tgt, T_gt


def loop_closure_pairs(n_pairs: int = 64, n_tgt: int = 200_000, rings: int = 32, azimuths: int = 1875,
                       first: int = 0, count: int | None = None):
    """Config 4: independent (scan, submap) pairs along the canyon; pair i is reproducible on its own.

    Yields (index, source, target, T_gt_i). Targets are local maps of a 120 m stretch around pose i,
    expressed in that pose's nominal frame.
    """
    scene = make_scene()
    d = math.pi / 180.0
    count = n_pairs - first if count is None else count
    for i in range(first, first + count):
        r = Rng(4000 + 10 * i)
        u = r.uniform(8)
        x_along = -100.0 + 200.0 * (i + 0.5) / n_pairs
        T_rel = pose_matrix((0.5 * (u[0] - 0.5), 0.4 * (u[1] - 0.5), 0.1 * (u[2] - 0.5)),
                            (0.6 * d * (u[3] - 0.5), 0.6 * d * (u[4] - 0.5), 3.0 * d * (u[5] - 0.5)))
        tgt = sample_map(scene, n_tgt, stream=4000 + 10 * i + 1, x_window=(x_along - 60.0, x_along + 60.0))
        tgt[:, 0] -= np.float32(x_along)
        src = make_scan(scene, rings, azimuths, sensor_pose(T_rel, x_along), stream=4000 + 10 * i + 2)
        yield i, src, tgt, T_rel


def pose_error(T_a: np.ndarray, T_b: np.ndarray):
    """(translation error [m], rotation angle of R_a R_b^T [rad]).

    The angle is taken from the skew part of R_a R_b^T (atan2 of |vee| and the trace term): acos((tr-1)/2) alone
    turns the 6e-8 quantisation of float32 matrix entries into ~4e-4 rad of fake error near the identity."""
    A = np.asarray(T_a, dtype=np.float64)
    B = np.asarray(T_b, dtype=np.float64)
    dt = float(np.linalg.norm(A[:3, 3] - B[:3, 3]))
    R = A[:3, :3] @ B[:3, :3].T
    v = 0.5 * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return dt, float(math.atan2(np.linalg.norm(v), 0.5 * (np.trace(R) - 1.0)))
