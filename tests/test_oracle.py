"""CPU tests that PIN THE ORACLE (no GPU): README known answers, analytic KATs, round trips.

The reference has no tests; its only known-answer numbers are the fitness values printed in
Thirdparty/ndt_omp_ros2/README.md:19-52 for the two vendored scans. The oracle must reproduce them before it is
trusted as the checker of the CUDA path.
"""
import math

import numpy as np
import pytest


def test_readme_fitness_known_answers(oracle_mod, golden):
    # apps/align.cpp:90-104: NDT resolution 1.0, identity guess, defaults (eps 0.1, 35 iterations, step 0.1)
    for name, method in (("DIRECT7", oracle_mod.DIRECT7), ("DIRECT1", oracle_mod.DIRECT1), ("KDTREE", oracle_mod.KDTREE)):
        n = oracle_mod.NDT(resolution=1.0, search_method=method)
        n.set_target(golden["target"])
        n.set_source(golden["source"])
        T = n.align()
        assert n.converged
        assert abs(n.fitness() - golden["readme_fitness"][name]) < 5e-6, name
        np.testing.assert_allclose(T, np.array(golden["ndt"][name]["final_transformation"]), atol=2e-6)
        assert n.iterations == golden["ndt"][name]["iterations"]


def test_readme_gicp_fitness_known_answer(oracle_mod, golden):
    """README.md:9-17 prints pcl::GICP 0.220382 / pclomp::GICP 0.220388 for the same two clouds (apps/align.cpp:81-87, default
    parameters). The restated GICP lands at 0.22035: agreement to 1.7e-4 relative — closer than this an iterative solver with
    convergence thresholds (and the testGradient ambiguity noted in oracle/gicp.hpp) cannot be pinned from a printed value, and
    about six times the gap between the two reference implementations themselves."""
    g = oracle_mod.GICP()
    g.set_target(golden["target"])
    g.set_source(golden["source"])
    g.align()
    assert g.converged
    assert abs(g.fitness() - 0.220388) < 1e-4
    assert abs(g.fitness() - 0.220382) < 1e-4


def test_oracle_thread_count_invariance(oracle_mod, golden):
    poses = []
    for nt in (1, 3, oracle_mod.max_threads()):
        n = oracle_mod.NDT(resolution=1.0, num_threads=nt)
        n.set_target(golden["target"])
        n.set_source(golden["source"])
        poses.append(n.align())
    for p in poses[1:]:
        np.testing.assert_allclose(p, poses[0], atol=1e-6)


def test_gauss_constants(oracle_mod):
    # res 1: d1=-2.217225, d2=0.433123; res 2: -4.196518, 0.248479; res 5: -6.931205, 0.149547
    for res, d1, d2 in ((1.0, -2.217225, 0.433123), (2.0, -4.196518, 0.248479), (5.0, -6.931205, 0.149547)):
        g = oracle_mod.NDT(resolution=res).gauss()
        assert abs(g[0] - d1) < 2e-6 and abs(g[1] - d2) < 2e-6


def test_euler_round_trip(oracle_mod):
    rng = np.random.default_rng(3)
    for _ in range(200):
        p = np.concatenate([rng.normal(size=3), rng.uniform(-1.2, 1.2, size=3)])
        T = oracle_mod.pose_to_matrix(p)
        ang = oracle_mod.euler_angles_012(T[:3, :3])
        assert 0.0 <= ang[0] <= math.pi + 1e-6  # Eigen convention: first angle in [0, pi]
        T2 = oracle_mod.pose_to_matrix(np.concatenate([p[:3], ang]))
        np.testing.assert_allclose(T2, T, atol=3e-6)
    assert np.all(oracle_mod.euler_angles_012(np.eye(3)) == 0)


def test_sym_eigen_and_inverse(oracle_mod):
    rng = np.random.default_rng(5)
    for _ in range(100):
        A = rng.normal(size=(3, 3))
        S = A @ A.T + 1e-3 * np.eye(3)
        ev, V = oracle_mod.sym_eigen3(S)
        assert ev[0] <= ev[1] <= ev[2]
        np.testing.assert_allclose(V @ np.diag(ev) @ V.T, S, atol=1e-12 * np.abs(S).max() + 1e-14)
        np.testing.assert_allclose(oracle_mod.mat3_inverse(S) @ S, np.eye(3), atol=1e-9)


def test_svd6_solve(oracle_mod):
    rng = np.random.default_rng(7)
    for _ in range(50):
        A = rng.normal(size=(6, 6))
        H = A + A.T
        b = rng.normal(size=6)
        np.testing.assert_allclose(oracle_mod.svd6_solve(H, b), np.linalg.solve(H, b), rtol=1e-8, atol=1e-10)
    # rank deficient: minimum-norm least squares like JacobiSVD::solve
    H = np.diag([4.0, 3.0, 2.0, 1.0, 0.0, 0.0])
    b = np.arange(1.0, 7.0)
    np.testing.assert_allclose(oracle_mod.svd6_solve(H, b), np.linalg.pinv(H) @ b, atol=1e-12)


def test_angle_tables_sy_quirk(oracle_mod):
    # ndt_omp_impl.hpp:381 keeps +sy in row d1 of the f32 table (the f64 vector at :359 has -sy)
    p = np.array([0, 0, 0, 0.3, -0.2, 0.5])
    j, h = oracle_mod.angle_tables(p)
    assert abs(h[6, 2] - math.sin(-0.2)) < 1e-7
    # below the 1e-4 snap threshold the tables are those of the identity rotation
    j0, h0 = oracle_mod.angle_tables(np.array([0, 0, 0, 5e-5, -5e-5, 9e-5]))
    jz, hz = oracle_mod.angle_tables(np.zeros(6))
    np.testing.assert_array_equal(j0, jz)
    np.testing.assert_array_equal(h0, hz)


def test_mt_trial_value_cases(oracle_mod):
    # case 1 (f_t > f_l): cubic/quadratic of f(a) = (a-1)^2 between a_l=0 and a_t=3 → minimiser at 1
    f = lambda a: (a - 1.0) ** 2
    g = lambda a: 2 * (a - 1.0)
    a = oracle_mod.mt_trial(0.0, f(0), g(0), 0.0, f(0), g(0), 3.0, f(3), g(3))
    assert abs(a - 1.0) < 1e-12
    # case 2 (f_t <= f_l, derivatives of opposite sign): secant/cubic both give the minimiser for a quadratic
    a = oracle_mod.mt_trial(0.0, f(0), g(0), 0.0, f(0), g(0), 1.5, f(1.5), g(1.5))
    assert abs(a - 1.0) < 1e-12
    # update: U1 (f_t > f_l) moves the upper end
    conv, v = oracle_mod.mt_update(0.0, 1.0, -2.0, 0.0, 1.0, -2.0, 3.0, 4.0, 4.0)
    assert not conv and v[3] == 3.0 and v[0] == 0.0


def test_voxelgrid_matches_bruteforce(oracle_mod, golden):
    raw = golden["raw"]
    leaf = 0.5
    out = oracle_mod.voxelgrid(raw, leaf)
    # independent numpy restatement of the leaf indexing + centroid
    inv = np.float32(1.0) / np.float32(leaf)
    mn, mx = raw[:, :3].min(0), raw[:, :3].max(0)
    min_b = np.floor(mn * inv).astype(np.int64)
    max_b = np.floor(mx * inv).astype(np.int64)
    div = max_b - min_b + 1
    ijk = (np.floor(raw[:, :3] * inv) - min_b.astype(np.float32)).astype(np.int64)
    idx = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    uniq, inv_idx = np.unique(idx, return_inverse=True)
    assert len(out) == len(uniq)
    sums = np.zeros((len(uniq), 4))
    np.add.at(sums, inv_idx, raw.astype(np.float64))
    cnt = np.bincount(inv_idx)
    np.testing.assert_allclose(out, sums / cnt[:, None], rtol=2e-5, atol=2e-5)


def test_voxel_covariance_identity_quirk(oracle_mod):
    # a leaf's cov_ accumulates on top of Identity (voxel_grid_covariance_omp.h:101): for n points the result is
    # ((n-1)/n) * (population_cov + I/n) before eigenvalue inflation
    rng = np.random.default_rng(11)
    pts = (rng.normal(size=(40, 3)) * np.array([0.3, 0.25, 0.2]) + np.array([0.5, 0.5, 0.5])).astype(np.float32)
    pts = pts[np.all((pts > 0.02) & (pts < 0.98), axis=1)]
    n = oracle_mod.NDT(resolution=1.0)
    n.set_target(pts)
    v = n.voxels()
    assert len(v["idx"]) == 1 and v["npts"][0] == len(pts)
    k = len(pts)
    p64 = pts.astype(np.float64)
    pop = np.cov(p64.T, bias=True)
    expect = (k - 1) / k * (pop + np.eye(3) / k)
    np.testing.assert_allclose(v["cov"][0], expect, atol=1e-9)
    np.testing.assert_allclose(v["icov"][0], np.linalg.inv(expect), rtol=1e-7)
    np.testing.assert_allclose(v["mean"][0], p64.mean(0), atol=1e-12)


def _single_voxel_problem(oracle_mod):
    """One 4 m voxel holding a Gaussian blob; source points stay well inside it, so DIRECT1 sees exactly one voxel
    per point and the score is smooth (the full NDT objective jumps when a point changes cell)."""
    rng = np.random.default_rng(19)
    tgt = (rng.normal(size=(400, 3)) * np.array([0.5, 0.4, 0.3]) + 2.0).astype(np.float32)
    tgt = tgt[np.all((tgt > 0.05) & (tgt < 3.95), axis=1)]
    src = (rng.uniform(1.6, 2.4, size=(300, 3))).astype(np.float32)
    n = oracle_mod.NDT(resolution=4.0, search_method=oracle_mod.DIRECT1)
    n.set_target(tgt)
    n.set_source(src)
    assert len(n.voxels()["idx"]) == 1
    return n


def test_derivatives_finite_difference(oracle_mod):
    n = _single_voxel_problem(oracle_mod)
    p = np.array([0.03, -0.02, 0.01, 0.02, -0.03, 0.04])
    s0, g, H = n.derivatives(oracle_mod.pose_to_matrix(p), p)
    h = 1e-3
    G = np.zeros((6, 6))
    for k in range(6):
        d = np.zeros(6)
        d[k] = h
        sp, gp, _ = n.derivatives(oracle_mod.pose_to_matrix(p + d), p + d)
        sm, gm, _ = n.derivatives(oracle_mod.pose_to_matrix(p - d), p - d)
        assert abs((sp - sm) / (2 * h) - g[k]) <= 2e-2 * max(1.0, np.abs(g).max()), k
        G[:, k] = (gp - gm) / (2 * h)
    scale = np.abs(H).max()
    # the live f32 table carries +sy in d1 (ndt_omp_impl.hpp:381) where the true second derivative has -sy:
    # only H(4,4) may deviate from the finite-difference Hessian
    mask = np.ones((6, 6), dtype=bool)
    mask[4, 4] = False
    assert np.abs(G - H)[mask].max() <= 2e-2 * scale
    np.testing.assert_allclose(H, H.T, atol=1e-4 * scale)
    # the f64 radius-neighbourhood Hessian (computeHessian, impl.hpp:538-629) uses the correct sign
    H64 = n.hessian_radius(oracle_mod.pose_to_matrix(p), p)
    assert np.abs(G - H64).max() <= 2e-2 * scale


def test_align_recovers_known_transform(oracle_mod, pair_small):
    # A scan ray-cast from T_gt must register back to T_gt; the street canyon leaves x weakly observed, so the
    # bound is loose — parity, not accuracy, is what the engine is judged on.
    from lidarslam_ros2_b200 import synth

    src, tgt, T_gt = pair_small
    n = oracle_mod.NDT(resolution=2.0, transformation_epsilon=0.01)
    n.set_target(tgt)
    n.set_source(src)
    T = n.align()
    dt, dr = synth.pose_error(T, T_gt)
    assert n.converged and dt < 0.35 and dr < 0.01


def test_exact_nn_against_bruteforce(oracle_mod):
    rng = np.random.default_rng(13)
    t = rng.uniform(-5, 5, size=(3000, 3)).astype(np.float32)
    q = rng.uniform(-6, 6, size=(500, 3)).astype(np.float32)
    idx, d2 = oracle_mod.nn1(t, q)
    d = ((q[:, None, :].astype(np.float64) - t[None, :, :]) ** 2).sum(-1)
    np.testing.assert_array_equal(idx, d.argmin(1))
    np.testing.assert_allclose(d2, d.min(1), rtol=1e-5)


def test_gicp_oracle_recovers_transform(oracle_mod):
    from lidarslam_ros2_b200 import synth

    rng = np.random.default_rng(17)
    # smooth random surface patches
    u = rng.uniform(-3, 3, size=(4000, 2))
    tgt = np.stack([u[:, 0], u[:, 1], 0.3 * np.sin(u[:, 0]) + 0.2 * np.cos(1.7 * u[:, 1])], axis=1).astype(np.float32)
    T_gt = synth.pose_matrix((0.08, -0.05, 0.03), (0.01, -0.015, 0.02))
    Ti = np.linalg.inv(T_gt)
    src = (tgt[::2].astype(np.float64) @ Ti[:3, :3].T + Ti[:3, 3]).astype(np.float32)
    g = oracle_mod.GICP()
    g.set_target(tgt)
    g.set_source(src)
    T = g.align()
    dt, dr = synth.pose_error(T, T_gt)
    assert g.converged and dt < 5e-3 and dr < 5e-3
