"""include/sag_detmath.h vs glibc (through numpy): the deterministic routines are as accurate as libm."""
import numpy as np

import oracle as O


def _ulps(a, b):
    return np.abs(a - b) / np.spacing(np.abs(b))


def test_sincos_within_1ulp():
    rng = np.random.RandomState(0)
    x = np.concatenate([rng.uniform(-20, 20, 400000), rng.uniform(-1e4, 1e4, 200000), rng.normal(0, 1e-3, 100000),
                        np.arange(-64, 65) * (np.pi / 4)])
    s, c = O.detmath("sincos", x)
    assert np.abs(s - np.sin(x)).max() <= 1.2e-16 and np.abs(c - np.cos(x)).max() <= 1.2e-16
    assert (s * s + c * c - 1.0).__abs__().max() < 5e-16


def test_atan2_within_2ulp_and_special_cases():
    rng = np.random.RandomState(1)
    y, x = rng.normal(0, 2, 500000), rng.normal(0, 2, 500000)
    a = O.detmath("atan2", y, x)
    assert _ulps(a, np.arctan2(y, x)).max() <= 2.0
    ys = np.array([0.0, 0.0, 1.0, -1.0, 0.0, 1e-300, -1e-300])
    xs = np.array([1.0, -1.0, 0.0, 0.0, 0.0, -1.0, -1.0])
    np.testing.assert_allclose(O.detmath("atan2", ys, xs), np.arctan2(ys, xs), rtol=0, atol=1e-300)


def test_log_within_1ulp():
    rng = np.random.RandomState(2)
    u = 1.0 - rng.uniform(size=500000)
    u[:8] = [1.0, 0.5, 2.0 ** -53, 1e-10, 0.999999999, 0.7071067811865475, 0.7071067811865476, 0.25]
    l = O.detmath("log", u)
    ref = np.log(u)
    m = ref != 0
    assert _ulps(l[m], ref[m]).max() <= 1.0 and l[0] == 0.0


def test_folded_lidar_equals_literal_lidar():
    """sag_lidar_bin16 (one division, bin index from folds) vs the literal atan2 form of safe_adaptation_gym.py:204-223:
    random scenes, objects exactly on bin edges (every multiple of pi/8), on the axes, and at the robot's own position"""
    import oracle as O
    rng = np.random.RandomState(5)
    worst = 0.0
    for i in range(6000):
        n = rng.randint(1, 14)
        rx, ry = rng.uniform(-2, 2, 2)
        yaw = rng.uniform(-10, 10)
        xs, ys = rng.uniform(-3.5, 3.5, n), rng.uniform(-3.5, 3.5, n)
        if i % 5 == 0:   # bin edges in the ego frame
            yaw = 0.0
            k, d = rng.randint(0, 16, n), rng.uniform(0.05, 3.2, n)
            xs, ys = rx + d * np.cos(k * np.pi / 8), ry + d * np.sin(k * np.pi / 8)
        if i % 7 == 0:   # exactly on an axis / at the robot
            xs[0], ys[0] = rx, ry + (rng.randint(-1, 2)) * 0.5
        a, b = O.lidar(rx, ry, yaw, xs, ys), O.lidar(rx, ry, yaw, xs, ys, literal=True)
        worst = max(worst, float(np.abs(a - b).max()))
    assert worst < 1e-14, worst
