"""GPU parity tests of the GICP path (K5 kNN covariances, K6 correspondences, K7 cost/gradient + BFGS) against the CPU
oracle's restatement of pclomp::GeneralizedIterativeClosestPoint. Parity unpinned beyond the oracle (see
oracle/gicp.hpp): PCL's BFGS and FLANN are external; the pose tolerance is BASELINE.json's 1e-3 m / 1e-3 rad."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def b200():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the gpu tests must run on the B200 box (there is no CPU fallback)")
    import lidarslam_ros2_b200 as m

    return m


def _surface(n, seed):
    rng = np.random.default_rng(seed)
    u = rng.uniform(-3, 3, size=(n, 2))
    z = 0.3 * np.sin(u[:, 0]) + 0.2 * np.cos(1.7 * u[:, 1])
    wall = rng.uniform(-3, 3, size=(n // 3, 2))
    a = np.stack([u[:, 0], u[:, 1], z], axis=1)
    b = np.stack([wall[:, 0], np.full(len(wall), 3.0) + 0.05 * np.sin(3 * wall[:, 0]), 1.5 + 0.5 * wall[:, 1]], axis=1)
    return np.concatenate([a, b]).astype(np.float32)


def _pair(seed=17):
    from lidarslam_ros2_b200 import synth

    tgt = _surface(6000, seed)
    T_gt = synth.pose_matrix((0.08, -0.05, 0.03), (0.01, -0.015, 0.02))
    Ti = np.linalg.inv(T_gt)
    src = (tgt[::2].astype(np.float64) @ Ti[:3, :3].T + Ti[:3, 3]).astype(np.float32)
    return src, tgt, T_gt


def test_gicp_covariances_parity(b200, oracle_mod):
    src, tgt, _ = _pair()
    g = b200.GeneralizedIterativeClosestPoint()
    g.setInputTarget(tgt)
    g.setInputSource(src)
    g.align()
    o = oracle_mod.GICP()
    o.set_target(tgt)
    o.set_source(src)
    o.align()
    for which in ("source", "target"):
        cg, co = g.covariances(which), o.covariances(which)
        assert cg.shape == co.shape
        err = np.abs(cg - co).max(axis=(1, 2))
        # exact kNN on both sides; a handful of points have a tie at the k-th neighbour or a nearly isotropic
        # neighbourhood (the smallest-variance direction is then ill-defined)
        assert np.mean(err < 1e-6) > 0.995, (which, np.mean(err < 1e-6))


def test_gicp_align_parity(b200, oracle_mod):
    from lidarslam_ros2_b200 import synth

    src, tgt, T_gt = _pair()
    g = b200.GeneralizedIterativeClosestPoint()
    g.setInputTarget(tgt)
    g.setInputSource(src)
    Tg = g.align()
    o = oracle_mod.GICP()
    o.set_target(tgt)
    o.set_source(src)
    To = o.align()
    dt, dr = synth.pose_error(Tg, To)
    assert dt < 1e-3 and dr < 1e-3, (dt, dr)
    assert g.hasConverged() and o.converged
    assert g.numCorrespondences() == o.num_correspondences()
    # and both recover the known transform
    dt, dr = synth.pose_error(Tg, T_gt)
    assert dt < 5e-3 and dr < 5e-3
    assert abs(g.getFitnessScore() - o.fitness()) <= 1e-3 * max(o.fitness(), 1e-6) + 1e-7
    # with a guess and the node's parameters (scanmatcher_component.cpp:116-120)
    guess = synth.pose_matrix((0.05, -0.02, 0.0), (0.0, -0.005, 0.01)).astype(np.float32)
    g.setMaxCorrespondenceDistance(5.0)
    g.setTransformationEpsilon(1e-8)
    o.set("max_correspondence_distance", 5.0)
    o.set("transformation_epsilon", 1e-8)
    g.setMaximumIterations(30)
    o.set("max_iterations", 30)
    dt, dr = synth.pose_error(g.align(guess), o.align(guess))
    assert dt < 1e-3 and dr < 1e-3, (dt, dr)


def test_gicp_on_lidar_scene(b200, oracle_mod, pair_tiny):
    from lidarslam_ros2_b200 import synth

    src, tgt, _ = pair_tiny
    g = b200.GeneralizedIterativeClosestPoint()
    g.setMaxCorrespondenceDistance(5.0)
    g.setInputTarget(tgt)
    g.setInputSource(src)
    o = oracle_mod.GICP(max_correspondence_distance=5.0)
    o.set_target(tgt)
    o.set_source(src)
    Tg, To = g.align(), o.align()
    dt, dr = synth.pose_error(Tg, To)
    # A sparse LiDAR scan has line-like 20-NN neighbourhoods (points of one ring): the covariance then has two nearly
    # equal small singular values and the direction that receives gicp_epsilon (last column of U, gicp_omp_impl.hpp:
    # 110-120) is decided by rounding noise / the SVD's sweep order — implementation-defined even between Eigen versions.
    # Only a loose agreement can be asserted here; the tight 1e-3 parity is asserted on well-conditioned surfaces above.
    assert g.hasConverged() == o.converged
    assert dt < 5e-2 and dr < 1e-2, (dt, dr)
    assert abs(g.getFitnessScore() - o.fitness()) <= 0.05 * o.fitness()
