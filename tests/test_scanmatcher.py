"""Frontend session: device-resident map maintenance + per-frame scan preparation, against the
CPU restatement oracle/scanmatcher.py of scanmatcher_component.cpp."""
import numpy as np
import pytest

import oracle
import oracle.scanmatcher as osm
from lidarslam_ros2_b200 import synth


def _sorted(c):
    c = np.asarray(c)
    return c[np.lexsort((c[:, 2], c[:, 1], c[:, 0]))]


# ---------------------------------------------------------------- CPU: the oracle itself
def test_pose_math_roundtrip():
    rng = np.random.default_rng(3)
    for _ in range(50):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        M = osm.pose_matrix([1.0, -2.0, 0.5], q)
        assert np.allclose(M[:3, :3] @ M[:3, :3].T, np.eye(3), atol=1e-12)
        q2 = osm.quat_from_matrix(M[:3, :3])
        assert min(np.abs(q2 - q).max(), np.abs(q2 + q).max()) < 1e-12
    # the largest-diagonal branches (trace <= 0)
    for q in ([1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0.7, 0.7, 0.1, 0.05]):
        q = np.array(q, dtype=float) / np.linalg.norm(q)
        q2 = osm.quat_from_matrix(osm.pose_matrix([0, 0, 0], q)[:3, :3])
        assert min(np.abs(q2 - q).max(), np.abs(q2 + q).max()) < 1e-12


def test_transforms_match_matrix_product():
    rng = np.random.default_rng(5)
    c = rng.normal(size=(200, 4)).astype(np.float32) * 20
    M = osm.pose_matrix([3.0, -1.0, 0.2], [0.01, -0.02, 0.3, 0.95] / np.linalg.norm([0.01, -0.02, 0.3, 0.95]))
    ref = (c[:, :3].astype(np.float64) @ M[:3, :3].T + M[:3, 3])
    assert np.abs(osm.transform_f64(c, M)[:, :3] - ref).max() < 1e-5
    assert np.abs(osm.transform_f32(c, M.astype(np.float32))[:, :3] - ref).max() < 1e-4
    assert np.array_equal(osm.transform_f64(c, M)[:, 3], c[:, 3])  # intensity is copied


def test_oracle_frontend_tracks_the_drive():
    sm = osm.ScanMatcher(ndt_resolution=2.0, vg_size_for_input=0.4, vg_size_for_map=0.3, num_targeted_cloud=4, num_threads=8)
    n_upd = 0
    for k, (scan, T_gt) in enumerate(synth.drive_stream(8, rings=16, azimuths=300, step=0.6)):
        pose, final, upd = sm.receive_cloud(scan)
        n_upd += int(upd)
        dt, dr = synth.pose_error(final, T_gt)
        assert dt < 0.5 and dr < 0.02, (k, dt, dr)  # a 16-ring scan against a map of a few sparse scans: bounded drift
    assert n_upd >= 2 and len(sm.submaps) == 1 + n_upd
    # targeted = newest scan + at most num_targeted_cloud-1 previous submaps
    assert len(sm.targeted) <= sum(len(c) for c, _, _ in sm.submaps[-4:])


# ---------------------------------------------------------------- GPU: parity through the C-ABI
@pytest.mark.gpu
@pytest.mark.parametrize("use_filter", [False, True])
def test_frontend_stream_parity(use_filter):
    from lidarslam_ros2_b200.scanmatcher import ScanMatcher

    kw = dict(ndt_resolution=2.0, vg_size_for_input=0.4, vg_size_for_map=0.3, num_targeted_cloud=3,
              use_min_max_filter=use_filter, scan_min_range=2.0, scan_max_range=60.0)
    g = ScanMatcher(**kw)
    o = osm.ScanMatcher(num_threads=oracle.max_threads(), **kw)
    n_upd = 0
    for k, (scan, T_gt) in enumerate(synth.drive_stream(10, rings=16, azimuths=400, step=0.6)):
        pg, Tg, ug = g.receiveCloud(scan)
        po, To, uo = o.receive_cloud(scan)
        assert ug == uo, k
        n_upd += int(ug)
        dt, dr = synth.pose_error(Tg, To)
        assert dt < 1e-3 and dr < 1e-3, (k, dt, dr)  # north_star tolerance
        assert np.abs(pg - po).max() < 1e-3
        fs = g.filteredScan()
        assert len(fs) == len(o.filtered)
        assert np.abs(_sorted(fs) - _sorted(o.filtered)).max() < 1e-4
    assert n_upd >= 2
    assert g.numSubmaps() == len(o.submaps)
    # the device-resident targeted cloud: same points (VoxelGrid centroids agree to float rounding; order is defined)
    tg, to = g.targetedCloud(), o.targeted
    assert tg.shape == to.shape
    # poses of GPU and CPU differ by <= 1e-3 m, and so do the transformed newest-scan points
    assert np.abs(tg - to).max() < 5e-3
    for i in range(g.numSubmaps()):
        c, M, dist = g.submap(i)
        co, Mo, disto = o.submaps[i]
        assert c.shape == co.shape and np.abs(c - co).max() < 1e-4
        assert np.abs(M - Mo).max() < 2e-3 and abs(dist - disto) < 2e-3


@pytest.mark.gpu
def test_update_map_bitwise_given_the_same_inputs():
    """updateMap alone, fed identical poses: transform arithmetic is bit-exact, VoxelGrid centroids to rounding."""
    from lidarslam_ros2_b200.scanmatcher import ScanMatcher

    g = ScanMatcher(ndt_resolution=2.0, vg_size_for_input=0.4, vg_size_for_map=0.25, num_targeted_cloud=3)
    o = osm.ScanMatcher(ndt_resolution=2.0, vg_size_for_input=0.4, vg_size_for_map=0.25, num_targeted_cloud=3, num_threads=4)
    rng = np.random.default_rng(11)
    for k, (scan, T_gt) in enumerate(synth.drive_stream(5, rings=16, azimuths=300, step=1.0)):
        q = osm.quat_from_matrix(T_gt[:3, :3])
        pos = T_gt[:3, 3] + 1e-3 * rng.normal(size=3)
        final = osm.pose_matrix(pos, q).astype(np.float32)
        cloud = np.concatenate([scan, rng.uniform(0, 255, size=(len(scan), 1)).astype(np.float32)], axis=1)
        g.setScan(cloud)
        g.updateMap(final, pos, q, adopt_now=True)
        o.update_map(cloud, final, pos, q)
        tg, to = g.targetedCloud(), o.targeted
        assert tg.shape == to.shape
        m = len(o.submaps[-1][0])
        # previous submaps were filtered identically on both sides only up to float rounding of the centroids, so
        # compare the transformed points with a rounding-level tolerance, and the structure (counts, order) exactly
        assert np.abs(tg - to).max() < 2e-4, k
        assert np.array_equal(np.isfinite(tg), np.isfinite(to))
        c, M, _ = g.submap(k)
        assert len(c) == m and np.array_equal(M, o.submaps[k][1])


# ---------------------------------------------------------------- loop search (backend node, gbs.cpp:144-258)
def _out_and_back(n_out=5, step=2.0, rings=16, azimuths=300):
    """Scans at ground-truth poses: n_out submaps down the canyon, then back over the same ground (0.3 m to the side)."""
    scene = synth.make_scene()
    xs = [step * k for k in range(n_out)] + [step * k for k in range(n_out - 2, -1, -1)]
    ys = [0.0] * n_out + [0.3] * (n_out - 1)
    S0 = synth.sensor_pose(synth.pose_matrix((0, 0, 0), (0, 0, 0)), -40.0)
    for k, (x, y) in enumerate(zip(xs, ys)):
        Sk = synth.sensor_pose(synth.pose_matrix((x, y, 0.0), (0.0, 0.0, 0.01 * k)), -40.0)
        yield synth.make_scan(scene, rings, azimuths, Sk, stream=8800 + k), np.linalg.inv(S0) @ Sk


def test_oracle_search_loop_finds_the_revisit():
    o = osm.ScanMatcher(ndt_resolution=2.0, vg_size_for_input=0.4, vg_size_for_map=0.3, num_targeted_cloud=3, num_threads=8)
    for scan, T in _out_and_back():
        o.update_map_external(scan, T.astype(np.float32), T[:3, 3], osm.quat_from_matrix(T[:3, :3]))
    assert len(o.submaps) == 9 and abs(o.submaps[-1][2] - 16.0) < 0.5
    reg = oracle.NDT(resolution=2.0, transformation_epsilon=0.01, max_iterations=100, search_method=oracle.DIRECT7, num_threads=8)
    far = o.search_loop(reg, voxel_leaf_size=0.3, distance_loop_closure=50.0, range_of_searching_loop_closure=1.0, search_submap_num=1)
    assert not far["is_candidate"] and far["id_min"] == -1  # never 50 m apart along the path
    r = o.search_loop(reg, voxel_leaf_size=0.3, distance_loop_closure=5.0, range_of_searching_loop_closure=1.0, search_submap_num=1)
    assert r["is_candidate"] and r["id_min"] == 0 and r["accepted"], r
    assert r["fitness"] < 1.0 and abs(r["min_dist"] - 0.3) < 1e-6
    # poses are ground truth, so the registration correction is small and the edge is the true relative pose
    M0, ML = o.submaps[0][1], o.submaps[-1][1]
    dt, dr = synth.pose_error(r["relative_pose"], np.linalg.inv(M0) @ ML)
    assert dt < 0.3 and dr < 0.02, (dt, dr)


@pytest.mark.gpu
def test_search_loop_parity():
    from lidarslam_ros2_b200.scanmatcher import ScanMatcher, backend_registration

    kw = dict(ndt_resolution=2.0, vg_size_for_input=0.4, vg_size_for_map=0.3, num_targeted_cloud=3)
    g = ScanMatcher(**kw)
    o = osm.ScanMatcher(num_threads=oracle.max_threads(), **kw)
    for scan, T in _out_and_back():
        q = osm.quat_from_matrix(T[:3, :3])
        g.setScan(scan)
        g.updateMap(T.astype(np.float32), T[:3, 3], q, adopt_now=False)
        o.update_map_external(scan, T.astype(np.float32), T[:3, 3], q)
    greg = backend_registration("NDT", ndt_resolution=2.0)
    oreg = oracle.NDT(resolution=2.0, transformation_epsilon=0.01, max_iterations=100, search_method=oracle.DIRECT7,
                      num_threads=oracle.max_threads())
    args = dict(voxel_leaf_size=0.3, distance_loop_closure=5.0, range_of_searching_loop_closure=1.0, search_submap_num=1)
    rg, ro = g.searchLoop(greg, **args), o.search_loop(oreg, **args)
    assert rg["is_candidate"] == ro["is_candidate"] and rg["id_min"] == ro["id_min"] == 0
    assert rg["accepted"] == ro["accepted"] is True
    assert rg["n_source"] == ro["n_source"]
    # the target is a VoxelGrid of transformed VoxelGrid centroids: a centroid that differs in its last float bit between the
    # two sides may fall into the neighbouring leaf
    assert abs(rg["n_target"] - ro["n_target"]) <= 4
    assert abs(rg["min_dist"] - ro["min_dist"]) < 1e-9
    dt, dr = synth.pose_error(rg["final"], ro["final"])
    assert dt < 1e-3 and dr < 1e-3, (dt, dr)
    assert abs(rg["fitness"] - ro["fitness"]) < 1e-3 * max(1.0, ro["fitness"])
    dt, dr = synth.pose_error(rg["relative_pose"], ro["relative_pose"])
    assert dt < 2e-3 and dr < 1e-3
    none = g.searchLoop(greg, voxel_leaf_size=0.3, distance_loop_closure=50.0, range_of_searching_loop_closure=1.0, search_submap_num=1)
    assert not none["is_candidate"] and none["id_min"] == -1


@pytest.mark.gpu
def test_search_loop_all_candidates():
    """b200sm_search_loop_all: every gated candidate, each registered exactly like searchLoop registers its single one. With a
    wide gate on the out-and-back drive several old submaps qualify; the closest one must reproduce searchLoop bitwise, the
    set must be the gate's set, and the shards of a 2-way split must reassemble the full list."""
    from lidarslam_ros2_b200.scanmatcher import ScanMatcher, backend_registration

    kw = dict(ndt_resolution=2.0, vg_size_for_input=0.4, vg_size_for_map=0.3, num_targeted_cloud=3)
    g = ScanMatcher(**kw)
    poses = []
    for scan, T in _out_and_back():
        g.setScan(scan)
        g.updateMap(T.astype(np.float32), T[:3, 3], osm.quat_from_matrix(T[:3, :3]), adopt_now=False)
        poses.append(T)
    greg = backend_registration("NDT", ndt_resolution=2.0)
    args = dict(voxel_leaf_size=0.3, distance_loop_closure=5.0, range_of_searching_loop_closure=4.5, search_submap_num=1)
    single = g.searchLoop(greg, **args)
    allc = g.searchLoopAll(greg, **args)
    assert len(allc) >= 2 and allc[0]["n_candidates_total"] == len(allc)
    ids = [c["id_min"] for c in allc]
    assert ids == sorted(ids)
    # the gate, recomputed here: travelled distance > 5 m behind the newest submap and position within 4.5 m
    latest = poses[-1][:3, 3]
    dist_along = np.concatenate([[0.0], np.cumsum([np.linalg.norm(poses[k][:3, 3] - poses[k - 1][:3, 3]) for k in range(1, len(poses))])])
    want = [k for k in range(len(poses)) if dist_along[-1] - dist_along[k] > 5.0 and np.linalg.norm(latest - poses[k][:3, 3]) < 4.5]
    assert ids == want, (ids, want)
    best = min(allc, key=lambda c: c["min_dist"])
    assert best["id_min"] == single["id_min"] and np.array_equal(best["final"], single["final"]) and best["fitness"] == single["fitness"]
    assert all(c["fitness"] > 0 and c["n_target"] > 0 for c in allc)
    shards = g.searchLoopAll(greg, shard_rank=0, shard_world=2, **args) + g.searchLoopAll(greg, shard_rank=1, shard_world=2, **args)
    shards.sort(key=lambda c: c["id_min"])
    assert [c["id_min"] for c in shards] == ids
    for a, b in zip(shards, allc):
        assert np.array_equal(a["final"], b["final"]) and a["fitness"] == b["fitness"]


@pytest.mark.gpu
def test_imported_submaps_give_the_same_loop_search():
    """A backend in its own process: submaps read back from the frontend session (the SubMap messages) and imported into a
    fresh session must give the identical loop search."""
    from lidarslam_ros2_b200.scanmatcher import ScanMatcher, backend_registration

    kw = dict(ndt_resolution=2.0, vg_size_for_input=0.4, vg_size_for_map=0.3, num_targeted_cloud=3)
    g = ScanMatcher(**kw)
    for scan, T in _out_and_back():
        g.setScan(scan)
        g.updateMap(T.astype(np.float32), T[:3, 3], osm.quat_from_matrix(T[:3, :3]), adopt_now=False)
    b = ScanMatcher(**kw)
    for i in range(g.numSubmaps()):
        cloud, M, dist = g.submap(i)
        b.importSubmap(cloud, M, dist)
    assert b.numSubmaps() == g.numSubmaps()
    reg = backend_registration("NDT", ndt_resolution=2.0)
    args = dict(voxel_leaf_size=0.3, distance_loop_closure=5.0, range_of_searching_loop_closure=1.0, search_submap_num=1)
    ra, rb = g.searchLoop(reg, **args), b.searchLoop(reg, **args)
    assert ra["id_min"] == rb["id_min"] and ra["accepted"] == rb["accepted"]
    assert np.array_equal(ra["final"], rb["final"]) and ra["fitness"] == rb["fitness"]
    assert np.array_equal(ra["relative_pose"], rb["relative_pose"])
