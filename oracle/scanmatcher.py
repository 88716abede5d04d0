"""ORACLE — TEST INFRASTRUCTURE ONLY. CPU restatement of the reference's frontend around align():
scanmatcher/src/scanmatcher_component.cpp — cloud_callback range filter (:211-219), initializeMap (:257-297),
receiveCloud (:299-389), publishMapAndPose (:391-434, with the map update applied synchronously, i.e. the mapping
thread always finishing before the next scan), updateMap (:438-491), getTransformation (:493-499).
numpy float32 / float64 elementwise arithmetic is IEEE and un-fused, so the association written here is exactly what is
computed. The registration, VoxelGrid and fitness pieces are the C++ oracle (oracle/ndt.hpp, voxelgrid.hpp).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np

import oracle


def transform_f32(cloud: np.ndarray, T: np.ndarray) -> np.ndarray:
    """pcl::transformPointCloud(in, out, Matrix4f): xyz <- R xyz + t in float, ((m0 x + m1 y) + m2 z) + m3 (the
    association oracle/ndt.hpp uses for the same PCL call); other fields copied."""
    T = np.asarray(T, dtype=np.float32)
    c = np.asarray(cloud, dtype=np.float32)
    out = c.copy()
    x, y, z = c[:, 0], c[:, 1], c[:, 2]
    for r in range(3):
        out[:, r] = ((T[r, 0] * x + T[r, 1] * y) + T[r, 2] * z) + T[r, 3]
    return out


def transform_f64(cloud: np.ndarray, M: np.ndarray) -> np.ndarray:
    """pcl::transformPointCloud(in, out, Matrix4d) (updateMap :459-462, submap_affine.matrix()): the generic
    Transformer<double>, static_cast<float>(m0 x + m1 y + m2 z + m3) left to right in double."""
    M = np.asarray(M, dtype=np.float64)
    c = np.asarray(cloud, dtype=np.float32)
    out = c.copy()
    x, y, z = c[:, 0].astype(np.float64), c[:, 1].astype(np.float64), c[:, 2].astype(np.float64)
    for r in range(3):
        out[:, r] = (((M[r, 0] * x + M[r, 1] * y) + M[r, 2] * z) + M[r, 3]).astype(np.float32)
    return out


def pose_matrix(position, quat_xyzw) -> np.ndarray:
    """tf2::fromMsg(pose, Affine3d) = Translation3d * Quaterniond(w, x, y, z) (Eigen toRotationMatrix), 4x4 double."""
    x, y, z, w = (float(v) for v in quat_xyzw)
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    M = np.eye(4)
    M[0, :3] = [1.0 - (tyy + tzz), txy - twz, txz + twy]
    M[1, :3] = [txy + twz, 1.0 - (txx + tzz), tyz - twx]
    M[2, :3] = [txz - twy, tyz + twx, 1.0 - (txx + tyy)]
    M[:3, 3] = position
    return M


def quat_from_matrix(R) -> np.ndarray:
    """Eigen::Quaterniond(Matrix3d) (publishMapAndPose :397-398); returns x, y, z, w."""
    m = np.asarray(R, dtype=np.float64)
    q = np.zeros(4)
    t = m[0, 0] + m[1, 1] + m[2, 2]
    if t > 0.0:
        t = np.sqrt(t + 1.0)
        q[3] = 0.5 * t
        t = 0.5 / t
        q[0] = (m[2, 1] - m[1, 2]) * t
        q[1] = (m[0, 2] - m[2, 0]) * t
        q[2] = (m[1, 0] - m[0, 1]) * t
    else:
        i = 0
        if m[1, 1] > m[0, 0]:
            i = 1
        if m[2, 2] > m[i, i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        t = np.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0)
        q[i] = 0.5 * t
        t = 0.5 / t
        q[3] = (m[k, j] - m[j, k]) * t
        q[j] = (m[j, i] + m[i, j]) * t
        q[k] = (m[k, i] + m[i, k]) * t
    return q


class ScanMatcher:
    def __init__(self, registration_method="NDT", ndt_resolution=5.0, gicp_corr_dist_threshold=5.0, trans_for_mapupdate=1.5,
                 vg_size_for_input=0.2, vg_size_for_map=0.1, use_min_max_filter=False, scan_min_range=0.1,
                 scan_max_range=100.0, num_targeted_cloud=10, num_threads=None):
        self.method = registration_method
        if registration_method == "NDT":  # :103-111
            self.reg = oracle.NDT(resolution=ndt_resolution, transformation_epsilon=0.01, search_method=oracle.DIRECT7,
                                  num_threads=num_threads or oracle.max_threads())
        else:  # :113-118
            self.reg = oracle.GICP(max_correspondence_distance=gicp_corr_dist_threshold, transformation_epsilon=1e-8)
        self.trans_for_mapupdate = trans_for_mapupdate
        self.vg_in, self.vg_map = vg_size_for_input, vg_size_for_map
        self.use_min_max, self.rmin, self.rmax = use_min_max_filter, scan_min_range, scan_max_range
        self.num_targeted_cloud = num_targeted_cloud
        self.position, self.quat = np.zeros(3), np.array([0.0, 0.0, 0.0, 1.0])
        self.previous_position = np.zeros(3)
        self.latest_distance, self.trans = 0.0, 0.0
        self.submaps = []  # (filtered sensor-frame cloud (m,4) f32, pose matrix 4x4 f64, distance)
        self.targeted = None
        self.pending = False
        self.initial = False
        self.filtered = None

    def set_initial_pose(self, position, quat_xyzw):
        self.position = np.asarray(position, dtype=np.float64).copy()
        self.previous_position = self.position.copy()
        self.quat = np.asarray(quat_xyzw, dtype=np.float64).copy()

    def _range_filter(self, cloud):
        if not self.use_min_max:
            return cloud
        r = np.sqrt(cloud[:, 0].astype(np.float64) ** 2 + cloud[:, 1].astype(np.float64) ** 2)  # :213-214
        return cloud[(self.rmin < r) & (r < self.rmax)]

    def update_map(self, cloud, final_T, position, quat):
        """updateMap :438-491 (initializeMap :257-297 when there is no submap yet)."""
        filtered = oracle.voxelgrid(cloud, self.vg_map)
        parts = [transform_f32(filtered, final_T)]
        n_sub = len(self.submaps)
        for i in range(self.num_targeted_cloud - 1):
            if n_sub - 1 - i < 0:
                continue
            c, M, _ = self.submaps[n_sub - 1 - i]
            parts.append(transform_f64(c, M))
        self.targeted = np.concatenate(parts, axis=0)
        self.submaps.append((filtered, pose_matrix(position, quat), self.latest_distance))
        self.pending = True

    def update_map_external(self, cloud, final_T, position, quat):
        """The caller-driven form (b200sm_update_map): advances latest_distance_ by the distance to the previous submap
        position (updateMap :471) and updates the map."""
        position = np.asarray(position, dtype=np.float64)
        if self.submaps:
            self.trans = float(np.sqrt(np.sum((position - self.previous_position) ** 2)))
            self.latest_distance += self.trans
        self.previous_position = position.copy()
        c = np.ascontiguousarray(cloud, dtype=np.float32)
        if c.shape[1] < 4:
            c = np.concatenate([c[:, :3], np.zeros((len(c), 1), dtype=np.float32)], axis=1)
        self.update_map(c, final_T, position, quat)

    def search_loop(self, reg, voxel_leaf_size=0.2, threshold_loop_closure_score=1.0, distance_loop_closure=20.0,
                    range_of_searching_loop_closure=20.0, search_submap_num=3):
        """GraphBasedSlamComponent::searchLoop (graph_based_slam_component.cpp:144-258) over this frontend's submaps.
        `reg` is the backend's registration oracle (gbs.cpp:47-64). Returns a dict like b200sm_loop_result."""
        out = {"is_candidate": False, "id_min": -1, "accepted": False}
        n_sub = len(self.submaps)
        if n_sub == 0:
            return out
        lc, lM, ldist = self.submaps[-1]
        min_dist, id_min = np.finfo(np.float64).max, 0
        for i, (c, M, dist_i) in enumerate(self.submaps):  # :187-204
            dist = float(np.sqrt(np.sum((lM[:3, 3] - M[:3, 3]) ** 2)))
            if ldist - dist_i > distance_loop_closure and dist < range_of_searching_loop_closure:
                out["is_candidate"] = True
                if dist < min_dist:
                    id_min, min_dist = i, dist
        if not out["is_candidate"]:
            return out
        out["id_min"], out["min_dist"] = id_min, min_dist
        src = transform_f32(lc, lM.astype(np.float32))  # :165-176
        parts = []
        for j in range(2 * search_submap_num + 1):  # :207-221 (indices past the newest submap: skipped, see b200reg.h)
            idx = id_min + j - search_submap_num
            if idx < 0 or idx >= n_sub:
                continue
            c, M, _ = self.submaps[idx]
            parts.append(transform_f32(c, M.astype(np.float32)))
        tgt = oracle.voxelgrid(np.concatenate(parts, axis=0), voxel_leaf_size)  # :223-225
        reg.set_source(src[:, :3])
        reg.set_target(tgt[:, :3])
        final = np.asarray(reg.align(), dtype=np.float32)  # :229
        fitness = reg.fitness()  # :230
        out.update(fitness=fitness, final=final, n_source=len(src), n_target=len(tgt))
        if fitness < threshold_loop_closure_score:  # :232-246
            out["accepted"] = True
            to = final.astype(np.float64) @ lM
            fr = self.submaps[id_min][1]
            inv = np.eye(4)
            inv[:3, :3] = fr[:3, :3].T
            inv[:3, 3] = -fr[:3, :3].T @ fr[:3, 3]
            out["relative_pose"] = inv @ to
        return out

    def _adopt(self, gicp_filter: bool):
        if not self.pending:
            return
        t = self.targeted
        if gicp_filter:
            t = oracle.voxelgrid(t, self.vg_in)  # :307-313
        self.reg.set_target(t[:, :3])
        self.pending = False

    def receive_cloud(self, points):
        cloud = np.ascontiguousarray(points, dtype=np.float32)
        if cloud.shape[1] < 4:
            cloud = np.concatenate([cloud[:, :3], np.zeros((len(cloud), 1), dtype=np.float32)], axis=1)
        cloud = self._range_filter(cloud)
        if not self.initial:
            self.initial = True
            sim = pose_matrix(self.position, self.quat).astype(np.float32)
            self.update_map(cloud, sim, self.position, self.quat)
            self._adopt(False)
        self._adopt(self.method == "GICP")
        self.filtered = oracle.voxelgrid(cloud, self.vg_in)  # :323-328
        self.reg.set_source(self.filtered[:, :3])
        sim = pose_matrix(self.position, self.quat).astype(np.float32)  # :330, :493-499
        final = np.asarray(self.reg.align(sim), dtype=np.float32)
        pos = final[:3, 3].astype(np.float64)
        self.quat = quat_from_matrix(final[:3, :3].astype(np.float64))
        self.position = pos
        self.trans = float(np.sqrt(np.sum((pos - self.previous_position) ** 2)))
        updated = False
        if self.trans >= self.trans_for_mapupdate:  # :420
            self.previous_position = pos.copy()
            self.latest_distance += self.trans
            self.update_map(cloud, final, self.position, self.quat)
            updated = True
        return np.concatenate([self.position, self.quat]), final, updated
