"""CPU tests of the solver's scalar control math (csrc/ndt_math.cuh compiled for the host) against the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hm():
    src = os.path.join(HERE, "hostmath", "hostmath.cpp")
    lib = os.path.join(HERE, "hostmath", "libhostmath.so")
    deps = [src, os.path.join(HERE, "..", "lidarslam_ros2_b200", "csrc", "ndt_math.cuh"),
            os.path.join(HERE, "..", "lidarslam_ros2_b200", "csrc", "angle_table_code.inc"),
            os.path.join(HERE, "..", "lidarslam_ros2_b200", "csrc", "bfgs6.hpp"), os.path.join(HERE, "..", "oracle", "bfgs.hpp")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", src, "-o", lib])
    L = C.CDLL(lib)
    L.hm_mt_trial.restype = C.c_double
    L.hm_gauss.argtypes = [C.c_double, C.c_float, C.c_void_p]
    L.hm_bfgs_compare.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p]
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_angle_tables_match_oracle_and_coded_form(hm, oracle_mod):
    rng = np.random.default_rng(1)
    for _ in range(200):
        p = np.concatenate([rng.normal(size=3), rng.uniform(-3.1, 3.1, size=3)])
        if rng.random() < 0.2:
            p[3 + rng.integers(3)] = 5e-5  # below the 1e-4 snap (ndt_omp_impl.hpp:292)
        j = np.zeros(24, dtype=np.float32)
        h = np.zeros(45, dtype=np.float32)
        jd = np.zeros(24)
        hd = np.zeros(45)
        hm.hm_angle_tables(_p(p), _p(j), _p(h), _p(jd), _p(hd))
        jo, ho = oracle_mod.angle_tables(p)
        np.testing.assert_array_equal(j.reshape(8, 3), jo)
        np.testing.assert_array_equal(h.reshape(15, 3), ho)
        coded = np.zeros(69)
        hm.hm_angle_tables_coded(_p(p), _p(coded))
        np.testing.assert_allclose(coded[:24], jd, rtol=0, atol=1e-15)
        np.testing.assert_allclose(coded[24:], hd, rtol=0, atol=1e-15)
        assert hd[20] == -h[20] or abs(hd[20] + h[20]) < 1e-7  # f64 -sy vs the live f32 +sy


def test_pose_and_euler_match_oracle(hm, oracle_mod):
    rng = np.random.default_rng(2)
    for _ in range(200):
        p = np.concatenate([rng.normal(size=3), rng.uniform(-1.3, 1.3, size=3)])
        T = np.zeros(12, dtype=np.float32)
        hm.hm_pose_to_matrix(_p(p), _p(T))
        To = oracle_mod.pose_to_matrix(p)
        np.testing.assert_array_equal(T.reshape(3, 4), To[:3])
        R = np.ascontiguousarray(To[:3, :3], dtype=np.float32).reshape(9)
        a = np.zeros(3, dtype=np.float32)
        hm.hm_euler_angles_012(_p(R), _p(a))
        np.testing.assert_array_equal(a, oracle_mod.euler_angles_012(To[:3, :3]))


def test_solve6_matches_svd(hm, oracle_mod):
    rng = np.random.default_rng(3)
    for k in range(100):
        A = rng.normal(size=(6, 6))
        H = (A + A.T) * np.array([1, 1, 1, 50, 50, 50])[:, None] * np.array([1, 1, 1, 50, 50, 50])[None, :]
        b = rng.normal(size=6)
        x = np.zeros(6)
        hm.hm_solve6(_p(np.ascontiguousarray(H)), _p(b), _p(x))
        np.testing.assert_allclose(x, oracle_mod.svd6_solve(H, b), rtol=1e-7, atol=1e-9)
    # rank-deficient → SVD fallback = minimum-norm solution
    H = np.diag([4.0, 3.0, 2.0, 1.0, 0.0, 0.0])
    b = np.arange(1.0, 7.0)
    x = np.zeros(6)
    hm.hm_solve6(_p(H), _p(b), _p(x))
    np.testing.assert_allclose(x, np.linalg.pinv(H) @ b, atol=1e-12)
    # zero matrix → zero step (the delta_p_norm == 0 exit, ndt_omp_impl.hpp:134)
    hm.hm_solve6(_p(np.zeros((6, 6))), _p(b), _p(x))
    assert np.all(x == 0)


def test_ldlt_solve6_matches_svd_on_definite_systems(hm, oracle_mod):
    """The controller's in-register Newton solve (ndt_math.cuh ldlt_solve6_upper): same x as JacobiSVD::solve on the
    negative-definite Hessians Newton works with; refuses (returns 0) collapsed pivots and non-finite input so that the
    pivoted-LU / SVD path takes over."""
    rng = np.random.default_rng(8)
    scale = np.array([1, 1, 1, 40, 40, 40.0])
    for k in range(200):
        A = rng.normal(size=(6, 6))
        H = -(A @ A.T + 0.05 * np.eye(6)) * scale[:, None] * scale[None, :]  # NDT's score Hessian is negative definite
        b = rng.normal(size=6) * scale
        x = np.zeros(6)
        assert hm.hm_ldlt_solve6(_p(np.ascontiguousarray(H)), _p(b), _p(x)) == 1
        ref = oracle_mod.svd6_solve(H, b)
        np.testing.assert_allclose(x, ref, rtol=1e-6 * np.linalg.cond(H) / 1e3 + 1e-9, atol=1e-10 * np.abs(ref).max())
    x = np.zeros(6)
    assert hm.hm_ldlt_solve6(_p(np.diag([4.0, 3.0, 2.0, 1.0, 0.0, 0.0])), _p(np.ones(6)), _p(x)) == 0  # rank deficient
    assert hm.hm_ldlt_solve6(_p(np.zeros((6, 6))), _p(np.ones(6)), _p(x)) == 0
    Hn = -np.eye(6)
    Hn[2, 4] = Hn[4, 2] = np.nan
    assert hm.hm_ldlt_solve6(_p(Hn), _p(np.ones(6)), _p(x)) == 0
    Hi = -np.eye(6)
    Hi[0, 0] = np.inf
    assert hm.hm_ldlt_solve6(_p(Hi), _p(np.ones(6)), _p(x)) == 0


def test_more_thuente_helpers_match_oracle(hm, oracle_mod):
    rng = np.random.default_rng(4)
    for _ in range(300):
        v = rng.normal(size=9)
        v[3] = v[0] + abs(v[3]) + 0.1
        v[6] = v[0] + 0.5 * (v[3] - v[0])
        try:
            a = hm.hm_mt_trial(_p(v))
            b = oracle_mod.mt_trial(*v)
        except Exception:
            continue
        assert (np.isnan(a) and np.isnan(b)) or a == b
        st = v[:6].copy()
        conv = hm.hm_mt_update(_p(st), _p(v[6:].copy()))
        conv_o, st_o = oracle_mod.mt_update(*v)
        assert bool(conv) == conv_o and np.array_equal(st, st_o)


def test_gauss_constants_match_oracle(hm, oracle_mod):
    for res in (0.5, 1.0, 2.0, 5.0):
        d = np.zeros(3)
        hm.hm_gauss(0.55, res, _p(d))
        np.testing.assert_array_equal(d, np.array(oracle_mod.NDT(resolution=res).gauss()))


def test_compact_sincos_within_one_ulp(hm):
    rng = np.random.default_rng(5)
    hm.hm_sincos_compact.argtypes = [C.c_double, C.c_void_p]
    xs = np.concatenate([rng.uniform(-3.2, 3.2, 20000), rng.uniform(-1e-3, 1e-3, 2000), rng.uniform(-40, 40, 2000),
                         np.array([0.0, np.pi / 2, -np.pi / 2, np.pi, -np.pi, np.pi / 4, 1e-5])])
    out = np.zeros(2)
    for x in xs:
        hm.hm_sincos_compact(float(x), _p(out))
        for got, ref in ((out[0], np.sin(x)), (out[1], np.cos(x))):
            assert abs(got - ref) <= 1.01 * np.spacing(abs(ref)) + 3e-17, (x, got, ref)


def test_templated_bfgs_matches_oracle_bfgs(hm):
    """csrc/bfgs6.hpp (host/device template, plain functor — the form the persistent GICP kernel instantiates) against
    oracle/bfgs.hpp on a smooth non-quadratic objective: same iterates, same number of functor evaluations."""
    rng = np.random.default_rng(21)
    for k in range(25):
        c = rng.normal(size=6)
        w = rng.uniform(0.5, 20.0, size=6)
        x0 = c + rng.normal(size=6) * 0.5
        out = np.zeros(22)
        hm.hm_bfgs_compare(_p(c), _p(w), _p(x0), 20, 1e-2, _p(out))
        np.testing.assert_array_equal(out[:6], out[6:12])
        np.testing.assert_array_equal(out[12:17], out[17:22])
        assert out[16] >= 1 and np.abs(out[:6] - c).max() < 0.5
