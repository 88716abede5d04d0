"""ORACLE — TEST INFRASTRUCTURE ONLY. PARITY UNPINNED (the reference has no test or known answer for this path).
CPU restatement of the reference's IMU de-skew, scanmatcher/include/scanmatcher/lidar_undistortion.hpp
(`LidarUndistortion::getImu` :52-106, `adjustDistortion` :110-226; used when `use_imu` is true,
scanmatcher_component.cpp:205-209 — off by default). SURVEY.md §8f row 4: not built on the GPU yet; this module is the
first half of that row (the oracle) plus the proof that the sequential state machine has an exact data-parallel form:

  * `adjust_distortion`           — the reference's loop, literally (half_passed flag, carried IMU ring pointer, `continue`).
  * `adjust_distortion_parallel`  — the same result from whole-array operations only: the `half_passed` switch is the first
    index at which a per-point predicate of the FIRST formula fires (a min-reduction); the carried pointer is a running
    maximum of independent lower bounds into the IMU ring (a prefix-max scan), re-evaluated until the set of skipped
    points is stable; everything else is per-point arithmetic. This is the formulation a CUDA kernel would use.

Float semantics: the reference mixes float members/locals with double literals; numpy float32/float64 scalars reproduce the
promotions written in the C++ (no FMA: the reference builds for baseline x86-64). Eigen's AngleAxisf product is restated as
the quaternion product it expands to; agreement with a real Eigen build is to float rounding, not bit-exact.
"""
from __future__ import annotations

import numpy as np

F = np.float32
QUE = 200  # imu_que_length_


def _quat_to_matrix_f(q):
    """Eigen::Quaternionf::toRotationMatrix (x, y, z, w), float."""
    x, y, z, w = (F(v) for v in q)
    tx, ty, tz = F(2) * x, F(2) * y, F(2) * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([[F(1) - (tyy + tzz), txy - twz, txz + twy],
                     [txy + twz, F(1) - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, F(1) - (txx + tyy)]], dtype=F)


def _quat_mul_f(a, b):
    """Eigen quaternion product, (x, y, z, w) float."""
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz], dtype=F)


def _rot_zyx_f(roll, pitch, yaw):
    """(AngleAxisf(yaw, Z) * AngleAxisf(pitch, Y) * AngleAxisf(roll, X)).toRotationMatrix() (:199-206): AngleAxis products
    are quaternion products in Eigen."""
    def aa(angle, axis):
        h = F(0.5) * F(angle)
        s, c = F(np.sin(h)), F(np.cos(h))
        q = np.zeros(4, dtype=F)
        q[axis] = s
        q[3] = c
        return q
    return _quat_to_matrix_f(_quat_mul_f(_quat_mul_f(aa(yaw, 2), aa(pitch, 1)), aa(roll, 0)))


class LidarUndistortion:
    def __init__(self, scan_period: float = 0.1):
        self.scan_period = float(scan_period)
        self.ptr_front, self.ptr_last, self.ptr_last_iter = 0, -1, 0
        self.time = np.zeros(QUE, dtype=np.float64)
        z = lambda: np.zeros(QUE, dtype=F)
        self.roll, self.pitch, self.yaw = z(), z(), z()
        self.velo = np.zeros((QUE, 3), dtype=F)
        self.shift = np.zeros((QUE, 3), dtype=F)
        self.ang_rot = np.zeros((QUE, 3), dtype=F)

    # ------------------------------------------------------------------ getImu :52-106
    def get_imu(self, angular_velo, acc, quat_xyzw, imu_time: float):
        R = _quat_to_matrix_f(quat_xyzw)
        roll = F(np.arctan2(R[2, 1], R[2, 2]))  # pcl::getEulerAngles
        pitch = F(np.arcsin(-R[2, 0]))
        yaw = F(np.arctan2(R[1, 0], R[0, 0]))
        self.ptr_last = (self.ptr_last + 1) % QUE
        if (self.ptr_last + 1) % QUE == self.ptr_front:
            self.ptr_front = (self.ptr_front + 1) % QUE
        k = self.ptr_last
        self.time[k] = imu_time
        self.roll[k], self.pitch[k], self.yaw[k] = roll, pitch, yaw
        a = (R @ np.asarray(acc, dtype=F)).astype(F)  # acc = rot * acc
        w = np.asarray(angular_velo, dtype=F)
        back = (k - 1 + QUE) % QUE
        dt = self.time[k] - self.time[back]  # double
        if dt < self.scan_period:
            # float member = float + float * double + float * double * double * 0.5  (evaluated in double, stored as float)
            self.shift[k] = (self.shift[back].astype(np.float64) + self.velo[back].astype(np.float64) * dt
                             + a.astype(np.float64) * dt * dt * 0.5).astype(F)
            self.velo[k] = (self.velo[back].astype(np.float64) + a.astype(np.float64) * dt).astype(F)
            self.ang_rot[k] = (self.ang_rot[back].astype(np.float64) + w.astype(np.float64) * dt).astype(F)

    # ------------------------------------------------------------------ shared scalar pieces
    def _orientation_range(self, cloud):
        start_ori = F(-np.arctan2(F(cloud[0, 1]), F(cloud[0, 0])))
        end_ori = F(-np.arctan2(F(cloud[-1, 1]), F(cloud[-1, 0])))
        if float(end_ori - start_ori) > 3 * np.pi:
            end_ori = F(float(end_ori) - 2 * np.pi)
        elif float(end_ori - start_ori) < np.pi:
            end_ori = F(float(end_ori) + 2 * np.pi)
        return start_ori, end_ori, F(end_ori - start_ori)

    def _interp(self, front, t):
        """rpy / shift / velo at time t given the ring pointer `front` (:169-197)."""
        if t > self.time[front]:
            return (np.array([self.roll[front], self.pitch[front], self.yaw[front]], dtype=F), self.shift[front].copy(),
                    self.velo[front].copy())
        back = (front - 1 + QUE) % QUE
        rf = F((t - self.time[back]) / (self.time[front] - self.time[back]))
        rb = F(1.0 - float(rf))
        rpy = np.array([self.roll[front] * rf + self.roll[back] * rb, self.pitch[front] * rf + self.pitch[back] * rb,
                        self.yaw[front] * rf + self.yaw[back] * rb], dtype=F)
        return rpy, (self.shift[front] * rf + self.shift[back] * rb).astype(F), (self.velo[front] * rf + self.velo[back] * rb).astype(F)

    # ------------------------------------------------------------------ adjustDistortion :110-226, literally
    def adjust_distortion(self, cloud: np.ndarray, scan_time: float) -> np.ndarray:
        out = np.array(cloud, dtype=F, copy=True)
        n = len(out)
        if n == 0:
            return out
        start_ori, end_ori, ori_diff = self._orientation_range(out)
        half_passed = False
        rpy_start = shift_start = velo_start = r_s_i = None
        for i in range(n):
            ori_h = F(-np.arctan2(out[i, 1], out[i, 0]))
            if not half_passed:
                if float(ori_h) < float(start_ori) - np.pi * 0.5:
                    ori_h = F(float(ori_h) + 2 * np.pi)
                elif float(ori_h) > float(start_ori) + np.pi * 1.5:
                    ori_h = F(float(ori_h) - 2 * np.pi)
                if float(F(ori_h - start_ori)) > np.pi:
                    half_passed = True
            else:
                ori_h = F(float(ori_h) + 2 * np.pi)
                if float(ori_h) < float(end_ori) - 1.5 * np.pi:
                    ori_h = F(float(ori_h) + 2 * np.pi)
                elif float(ori_h) > float(end_ori) + 0.5 * np.pi:
                    ori_h = F(float(ori_h) - 2 * np.pi)
            rel_time = F(float(F(F(ori_h - start_ori) / ori_diff)) * self.scan_period)
            if self.ptr_last > 0:
                front = self.ptr_last_iter
                t = scan_time + float(rel_time)
                while front != self.ptr_last:
                    if t < self.time[front]:
                        break
                    front = (front + 1) % QUE
                self.ptr_front = front
                if abs(t - self.time[front]) > self.scan_period:
                    continue  # (skips the pointer carry at the bottom of the loop, like the reference)
                rpy, shift, velo = self._interp(front, t)
                r_c = _rot_zyx_f(rpy[0], rpy[1], rpy[2])
                if i == 0:
                    rpy_start, shift_start, velo_start = rpy, shift, velo
                    r_s_i = r_c.T.copy()  # r_c.inverse() of a rotation (Eigen computes the general inverse; equal to rounding)
                elif r_s_i is not None:
                    sfs = (shift - shift_start - velo_start * rel_time).astype(F)
                    out[i, :3] = (r_s_i @ ((r_c @ out[i, :3]).astype(F) + sfs).astype(F)).astype(F)
            self.ptr_last_iter = self.ptr_front
        return out

    # ------------------------------------------------------------------ the same result without a sequential loop
    def adjust_distortion_parallel(self, cloud: np.ndarray, scan_time: float) -> np.ndarray:
        out = np.array(cloud, dtype=F, copy=True)
        n = len(out)
        if n == 0:
            return out
        start_ori, end_ori, ori_diff = self._orientation_range(out)
        ori = (-np.arctan2(out[:, 1], out[:, 0])).astype(F)
        so, eo = float(start_ori), float(end_ori)
        # formula A (before the half turn) for every point
        a = ori.copy()
        lo, hi = a.astype(np.float64) < so - np.pi * 0.5, a.astype(np.float64) > so + np.pi * 1.5
        a[lo] = (a[lo].astype(np.float64) + 2 * np.pi).astype(F)
        a[hi & ~lo] = (a[hi & ~lo].astype(np.float64) - 2 * np.pi).astype(F)
        fires = (a - start_ori).astype(F).astype(np.float64) > np.pi
        k = int(np.argmax(fires)) if fires.any() else n  # first index that sets half_passed (it still uses formula A itself)
        # formula B for every point after k
        b = (ori.astype(np.float64) + 2 * np.pi).astype(F)
        lo, hi = b.astype(np.float64) < eo - 1.5 * np.pi, b.astype(np.float64) > eo + 0.5 * np.pi
        b[lo] = (b[lo].astype(np.float64) + 2 * np.pi).astype(F)
        b[hi & ~lo] = (b[hi & ~lo].astype(np.float64) - 2 * np.pi).astype(F)
        ori_h = np.where(np.arange(n) <= k, a, b).astype(F)
        rel = (((ori_h - start_ori).astype(F) / ori_diff).astype(F).astype(np.float64) * self.scan_period).astype(F)
        if self.ptr_last <= 0:
            self.ptr_last_iter = self.ptr_front
            return out
        t = scan_time + rel.astype(np.float64)
        # ring positions in chronological order starting at the carried pointer
        base, last = self.ptr_last_iter, self.ptr_last
        span = (last - base) % QUE  # positions 0..span map to ring indices base..last
        ring = (base + np.arange(span + 1)) % QUE
        times = self.time[ring]
        # independent lower bound: first position whose time is > t (else the last position). The walk in the reference
        # stops at the first j with t < time[j]; times along the walk are non-decreasing for a live IMU stream.
        lb = np.minimum(np.searchsorted(times, t, side="right"), span)
        # carried pointer = running max over the previous NON-skipped points; iterate until the skipped set is stable
        skipped = np.zeros(n, dtype=bool)
        for _ in range(n + 1):
            contrib = np.where(skipped, -1, lb)
            carried = np.maximum.accumulate(np.concatenate([[0], contrib[:-1]]))
            front_pos = np.maximum(carried, lb)
            new_skipped = np.abs(t - times[front_pos]) > self.scan_period
            if np.array_equal(new_skipped, skipped):
                break
            skipped = new_skipped
        front = ring[front_pos]
        ok = ~skipped
        self.ptr_front = int(front[-1])  # assigned before the skip test, for every point
        if ok.any():
            self.ptr_last_iter = int(front[np.flatnonzero(ok)[-1]])  # carried only past non-skipped points
        if not ok[0]:
            return out  # the reference never initialises the start pose then: no point is corrected
        # per-point interpolation and correction (vectorised over the non-skipped points; point 0 defines the start pose)
        rpy0, shift0, velo0 = self._interp(int(front[0]), float(t[0]))
        r_s_i = _rot_zyx_f(rpy0[0], rpy0[1], rpy0[2]).T.copy()
        for i in np.flatnonzero(ok)[1:] if ok[0] else []:  # the arithmetic below is independent per point
            rpy, shift, velo = self._interp(int(front[i]), float(t[i]))
            r_c = _rot_zyx_f(rpy[0], rpy[1], rpy[2])
            sfs = (shift - shift0 - velo0 * rel[i]).astype(F)
            out[i, :3] = (r_s_i @ ((r_c @ out[i, :3]).astype(F) + sfs).astype(F)).astype(F)
        return out
