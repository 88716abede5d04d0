"""Host-side mirror of the reference's frontend node for the path around align():
`ScanMatcherComponent` (scanmatcher/src/scanmatcher_component.cpp) without ROS — cloud callback, initializeMap,
receiveCloud, publishMapAndPose, updateMap — on top of the `b200sm_*` session of the C-ABI (include/b200reg.h).
The submaps and the targeted cloud live on the GPU; a frame costs one host-to-device copy.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from .registration import B200RegError, GeneralizedIterativeClosestPoint, NormalDistributionsTransform, _as_cloud, _ptr


class ScanMatcher:
    """Parameters carry the reference's names and defaults (scanmatcher_component.cpp:26-50)."""

    def __init__(self, registration_method: str = "NDT", ndt_resolution: float = 5.0, ndt_num_threads: int = 0,
                 gicp_corr_dist_threshold: float = 5.0, trans_for_mapupdate: float = 1.5, vg_size_for_input: float = 0.2,
                 vg_size_for_map: float = 0.1, use_min_max_filter: bool = False, scan_min_range: float = 0.1,
                 scan_max_range: float = 100.0, num_targeted_cloud: int = 10, device: int = 0):
        self._lib = _capi.lib()
        if registration_method == "NDT":  # scanmatcher_component.cpp:97-107
            reg = NormalDistributionsTransform(device=device)
            reg.setResolution(ndt_resolution)
            reg.setTransformationEpsilon(0.01)
            reg.setNeighborhoodSearchMethod(2)  # pclomp::DIRECT7
            if ndt_num_threads > 0:
                reg.setNumThreads(ndt_num_threads)
        elif registration_method == "GICP":  # :108-115
            reg = GeneralizedIterativeClosestPoint(device=device)
            reg.setMaxCorrespondenceDistance(gicp_corr_dist_threshold)
            reg.setTransformationEpsilon(1e-8)
        else:
            raise ValueError("registration_method must be NDT or GICP")
        self.registration = reg
        h = C.c_void_p()
        rc = self._lib.b200sm_create(int(device), C.byref(h))
        if rc != 0:
            raise B200RegError(rc, "b200sm_create failed (no CUDA device? there is no CPU fallback)")
        self._h = h
        self._check(self._lib.b200sm_set_params(self._h, float(vg_size_for_input), float(vg_size_for_map), int(num_targeted_cloud),
                                                float(trans_for_mapupdate), int(bool(use_min_max_filter)),
                                                float(scan_min_range), float(scan_max_range)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.b200sm_destroy(h)
            self._h = None

    def _check(self, rc):
        if rc != 0:
            raise B200RegError(rc, self._lib.b200sm_last_error(self._h).decode())

    def setInitialPose(self, position, quat_xyzw):
        p = np.ascontiguousarray(position, dtype=np.float64)
        q = np.ascontiguousarray(quat_xyzw, dtype=np.float64)
        self._check(self._lib.b200sm_set_initial_pose(self._h, _ptr(p), _ptr(q)))

    def receiveCloud(self, points):
        """One frame (x, y, z[, intensity] rows). Returns (pose7 = position + quaternion xyzw, final 4x4, map_updated)."""
        p = _as_cloud(points)
        n, w = p.shape
        pose = np.zeros(7, dtype=np.float64)
        fin = np.zeros(16, dtype=np.float32)
        upd = C.c_int(0)
        self._check(self._lib.b200sm_receive_cloud(self._h, self.registration._h, _ptr(p), n, 4 * w, 12 if w >= 4 else -1,
                                                   _ptr(pose), _ptr(fin), C.byref(upd)))
        return pose, fin.reshape(4, 4).T.copy(), bool(upd.value)

    # ---- the pieces, for callers that drive the steps themselves ----
    def setScan(self, points) -> int:
        p = _as_cloud(points)
        n, w = p.shape
        m = C.c_size_t(0)
        self._check(self._lib.b200sm_set_scan(self._h, self.registration._h, _ptr(p), n, 4 * w, 12 if w >= 4 else -1, C.byref(m)))
        return int(m.value)

    def deskewNextScan(self, scan_time: float):
        """use_imu: de-skew the next frame on the device before the range filter (b200sm_deskew_next_scan)."""
        self._check(self._lib.b200sm_deskew_next_scan(self._h, float(scan_time)))

    def updateMap(self, final_transformation, position, quat_xyzw, adopt_now: bool = True):
        T = np.ascontiguousarray(np.asarray(final_transformation, dtype=np.float32).T).reshape(16)
        p = np.ascontiguousarray(position, dtype=np.float64)
        q = np.ascontiguousarray(quat_xyzw, dtype=np.float64)
        self._check(self._lib.b200sm_update_map(self._h, self.registration._h, _ptr(T), _ptr(p), _ptr(q), int(adopt_now)))

    def searchLoop(self, registration, voxel_leaf_size: float = 0.2, threshold_loop_closure_score: float = 1.0,
                   distance_loop_closure: float = 20.0, range_of_searching_loop_closure: float = 20.0,
                   search_submap_num: int = 3) -> dict:
        """GraphBasedSlamComponent::searchLoop (graph_based_slam_component.cpp:144-258) over the session's device-resident
        submaps; `registration` is the backend's engine (see backend_registration()). Parameter names and defaults are the
        node's (gbs.cpp:23-39)."""
        r = _capi.SmLoopResult()
        self._check(self._lib.b200sm_search_loop(self._h, registration._h, float(voxel_leaf_size), float(threshold_loop_closure_score),
                                                 float(distance_loop_closure), float(range_of_searching_loop_closure),
                                                 int(search_submap_num), C.byref(r)))
        out = {"is_candidate": bool(r.is_candidate), "id_min": int(r.id_min), "accepted": bool(r.accepted)}
        if r.is_candidate:
            out.update(min_dist=float(r.min_dist), fitness=float(r.fitness), n_source=int(r.n_source), n_target=int(r.n_target),
                       final=np.array(r.final_T, dtype=np.float32).reshape(4, 4).T.copy())
            if r.accepted:
                out["relative_pose"] = np.array(r.relative_pose, dtype=np.float64).reshape(4, 4).T.copy()
        return out

    def importSubmap(self, cloud, pose_matrix, distance: float):
        """Append a submap received as a message (filtered cloud, 4x4 pose, travelled distance): b200sm_import_submap."""
        p = _as_cloud(cloud)
        n, w = p.shape
        M = np.ascontiguousarray(np.asarray(pose_matrix, dtype=np.float64).T).reshape(16)
        self._check(self._lib.b200sm_import_submap(self._h, _ptr(p), n, 4 * w, 12 if w >= 4 else -1, _ptr(M), float(distance)))

    def searchLoopAll(self, registration, voxel_leaf_size: float = 0.2, threshold_loop_closure_score: float = 1.0,
                      distance_loop_closure: float = 20.0, range_of_searching_loop_closure: float = 20.0,
                      search_submap_num: int = 3, shard_rank: int = 0, shard_world: int = 1) -> list:
        """Every gated candidate instead of the closest one (b200sm_search_loop_all); one dict per candidate, ascending id."""
        cap = max(1, self.numSubmaps())
        arr = (_capi.SmLoopResult * cap)()
        n, tot = C.c_size_t(0), C.c_size_t(0)
        self._check(self._lib.b200sm_search_loop_all(self._h, registration._h, float(voxel_leaf_size), float(threshold_loop_closure_score),
                                                     float(distance_loop_closure), float(range_of_searching_loop_closure),
                                                     int(search_submap_num), int(shard_rank), int(shard_world), arr, cap,
                                                     C.byref(n), C.byref(tot)))
        out = []
        for k in range(n.value):
            r = arr[k]
            d = {"is_candidate": True, "id_min": int(r.id_min), "accepted": bool(r.accepted), "min_dist": float(r.min_dist),
                 "fitness": float(r.fitness), "n_source": int(r.n_source), "n_target": int(r.n_target),
                 "final": np.array(r.final_T, dtype=np.float32).reshape(4, 4).T.copy(), "n_candidates_total": int(tot.value)}
            if r.accepted:
                d["relative_pose"] = np.array(r.relative_pose, dtype=np.float64).reshape(4, 4).T.copy()
            out.append(d)
        return out

    # ---- read-back ----
    def stats(self) -> dict:
        st = _capi.SmStats()
        self._check(self._lib.b200sm_get_stats(self._h, C.byref(st)))
        return {k: getattr(st, k) for k, _ in _capi.SmStats._fields_}

    def numSubmaps(self) -> int:
        v = C.c_size_t(0)
        self._check(self._lib.b200sm_num_submaps(self._h, C.byref(v)))
        return int(v.value)

    def targetedCloud(self) -> np.ndarray:
        n = C.c_size_t(0)
        self._check(self._lib.b200sm_get_targeted(self._h, None, 0, C.byref(n)))
        out = np.empty((n.value, 4), dtype=np.float32)
        self._check(self._lib.b200sm_get_targeted(self._h, _ptr(out), n.value, C.byref(n)))
        return out

    def filteredScan(self) -> np.ndarray:
        n = C.c_size_t(0)
        self._check(self._lib.b200sm_get_filtered_scan(self._h, None, 0, C.byref(n)))
        out = np.empty((n.value, 4), dtype=np.float32)
        self._check(self._lib.b200sm_get_filtered_scan(self._h, _ptr(out), n.value, C.byref(n)))
        return out

    def submap(self, index: int):
        n = C.c_size_t(0)
        pose = np.zeros(16, dtype=np.float64)
        dist = C.c_double(0)
        self._check(self._lib.b200sm_get_submap(self._h, index, None, 0, C.byref(n), _ptr(pose), C.byref(dist)))
        out = np.empty((n.value, 4), dtype=np.float32)
        self._check(self._lib.b200sm_get_submap(self._h, index, _ptr(out), n.value, C.byref(n), _ptr(pose), C.byref(dist)))
        return out, pose.reshape(4, 4).T.copy(), float(dist.value)


def backend_registration(registration_method: str = "NDT", ndt_resolution: float = 5.0, ndt_num_threads: int = 0, device: int = 0):
    """The backend node's registration object (graph_based_slam_component.cpp:44-70)."""
    if registration_method == "NDT":
        reg = NormalDistributionsTransform(device=device)
        reg.setMaximumIterations(100)
        reg.setResolution(ndt_resolution)
        reg.setTransformationEpsilon(0.01)
        reg.setNeighborhoodSearchMethod(2)  # pclomp::DIRECT7
        if ndt_num_threads > 0:
            reg.setNumThreads(ndt_num_threads)
        return reg
    if registration_method == "GICP":
        reg = GeneralizedIterativeClosestPoint(device=device)
        reg.setMaxCorrespondenceDistance(30)
        reg.setMaximumIterations(100)
        reg.setTransformationEpsilon(1e-8)
        reg.setEuclideanFitnessEpsilon(1e-6)
        reg.setRANSACIterations(0)
        return reg
    raise ValueError("registration_method must be NDT or GICP")


class LidarUndistortion:
    """scanmatcher/include/scanmatcher/lidar_undistortion.hpp on the GPU session (b200sm_imu_*): getImu keeps the IMU ring
    on the host like the reference, adjustDistortion runs as CUDA kernels on the uploaded scan."""

    def __init__(self, device: int = 0, scan_period: float = 0.1, session=None):
        import ctypes as C

        from . import _capi

        self._C, self._lib = C, _capi.lib()
        self._own = session is None
        if session is None:
            h = C.c_void_p()
            rc = self._lib.b200sm_create(int(device), C.byref(h))
            if rc != 0:
                raise RuntimeError(f"b200sm_create failed ({rc}): no CUDA device? there is no CPU fallback")
            self._s = h
        else:
            self._s = session
        self.setScanPeriod(scan_period)

    def __del__(self):
        if getattr(self, "_own", False) and getattr(self, "_s", None):
            self._lib.b200sm_destroy(self._s)
            self._s = None

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"b200sm error {rc}: {self._lib.b200sm_last_error(self._s).decode()}")

    def setScanPeriod(self, scan_period: float):
        self._check(self._lib.b200sm_imu_set_scan_period(self._s, float(scan_period)))

    def getImu(self, angular_velo, acc, quat_xyzw, imu_time: float):
        a = np.ascontiguousarray(angular_velo, dtype=np.float32)
        b = np.ascontiguousarray(acc, dtype=np.float32)
        q = np.ascontiguousarray(quat_xyzw, dtype=np.float32)
        self._check(self._lib.b200sm_imu_push(self._s, a.ctypes.data, b.ctypes.data, q.ctypes.data, float(imu_time)))

    def adjustDistortion(self, cloud, scan_time: float) -> np.ndarray:
        """cloud: (N, >=3) float32 in firing order; returns the corrected copy (other columns untouched)."""
        c = np.array(cloud, dtype=np.float32, copy=True, order="C")
        ioff = 12 if c.shape[1] >= 4 else -1
        self._check(self._lib.b200sm_imu_adjust_distortion(self._s, c.ctypes.data, len(c), c.strides[0], ioff, float(scan_time)))
        return c

    def pointers(self):
        C = self._C
        a, b, c = C.c_int(0), C.c_int(0), C.c_int(0)
        self._check(self._lib.b200sm_imu_get_state(self._s, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def sample(self, index: int):
        C = self._C
        t = C.c_double(0)
        rpy, sh, ve = (np.zeros(3, dtype=np.float32) for _ in range(3))
        self._check(self._lib.b200sm_imu_get_sample(self._s, int(index), C.byref(t), rpy.ctypes.data, sh.ctypes.data, ve.ctypes.data))
        return t.value, rpy, sh, ve
