"""Size-independent properties of the oracle's force / step functions (CPU only).

The reference's testbenches pin no outputs (tests/test_oracle_kat.py), so besides the hand-computable
stimuli the oracle is held to what the formula of S/dxy.vhd:94-122, S/dzsoft.vhd:177-202, S/dxyz_soft.vhd:149-150,
S/cube.vhd:66-70 and S/fxyz.vhd:101-127 implies for ANY input: antisymmetric pair terms (sum of accelerations = 0
with unit masses), covariance under permutation / translation / scaling, the composite loop equal to the per-entity
functions applied pair by pair, and the kick-then-drift split of one step.  The same properties are asked of the GPU
path at full size in tests/test_gpu_parity.py (momentum at C2/C4/C5, permutation invariance, split independence).
"""
import ctypes as C

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

f32 = np.float32
SET = dict(max_examples=25, deadline=None, derandomize=True)     # same examples on every run


def _bodies(orc, n, seed):
    return orc.randomize(n, seed)


@settings(**SET)
@given(n=st.integers(2, 300), seed=st.integers(0, 2**31))
def test_momentum_is_conserved(orc, n, seed):
    # F_ij = -F_ji (d = target - this, S/dxy.vhd:95,98) and unit masses: the accelerations sum to zero
    b = orc.widen(_bodies(orc, n, seed))
    a = orc.accel_f64(b)
    assert np.abs(a.sum(axis=0)).max() <= 1e-12 * np.abs(a).sum(axis=0).max()


@settings(**SET)
@given(n=st.integers(2, 200), seed=st.integers(0, 2**31))
def test_permutation_covariance(orc, n, seed):
    b = orc.widen(_bodies(orc, n, seed))
    perm = np.random.default_rng(seed).permutation(n)
    a, ap = orc.accel_f64(b), orc.accel_f64(b[perm])
    assert orc.rel_err(ap, a[perm]).max() <= 1e-12
    # the FP32 sequential-j loop is order dependent, but only at rounding level
    b32 = _bodies(orc, n, seed)
    assert orc.rel_err(orc.accel_f32(b32[perm]), orc.accel_f32(b32)[perm]).max() <= 1e-4


@settings(**SET)
@given(n=st.integers(2, 200), seed=st.integers(0, 2**31), shift=st.floats(-4.0, 4.0))
def test_translation_invariance(orc, n, seed, shift):
    b = orc.widen(_bodies(orc, n, seed))
    c = b.copy()
    for k, s in zip("xyz", (shift, -0.5 * shift, 0.25 * shift)):
        c[k] = b[k] + s
    # differences of shifted doubles carry ~1e-16 * |shift| / |d| relative error; pairs closer than 1e-4 are rare at n <= 200
    assert orc.rel_err(orc.accel_f64(c), orc.accel_f64(b)).max() <= 1e-9


@settings(**SET)
@given(n=st.integers(2, 200), seed=st.integers(0, 2**31), lam=st.sampled_from([0.5, 2.0, 4.0, 0.125]))
def test_scaling_law(orc, n, seed, lam):
    # x -> lam x, eps -> lam^2 eps  =>  a -> a / lam^2 (power-of-two factors: exact in binary arithmetic)
    b = orc.widen(_bodies(orc, n, seed))
    c = b.copy()
    for k in "xyz":
        c[k] = lam * b[k]
    a = orc.accel_f64(b)
    with orc.softening(1e-9 * lam * lam):
        al = orc.accel_f64(c)
    np.testing.assert_allclose(al * lam * lam, a, rtol=1e-13, atol=0)


@settings(**SET)
@given(seed=st.integers(0, 2**31))
def test_composite_loop_equals_the_entities_pair_by_pair(orc, seed):
    # oracle_accel_f32 on a small system == dxy / dzsoft / dxyz_soft / rsqrt / cube / three FMAs applied in j order
    # (the rounding points of the reference pipeline), bit for bit
    n = 7
    b = _bodies(orc, n, seed)
    L = orc.load()
    got = orc.accel_f32(b)
    for i in range(n):
        acc = [f32(0), f32(0), f32(0)]
        this_ = np.array([b["x"][i], b["y"][i], b["z"][i]], dtype=f32)
        for j in range(n):
            tgt = np.array([b["x"][j], b["y"][j], b["z"][j]], dtype=f32)
            d = np.zeros(3, dtype=f32)
            s = L.oracle_dxyz_soft(orc._p(this_), orc._p(tgt), orc._p(d))
            w = L.oracle_cube(L.oracle_rsqrt(s))
            for k in range(3):                                   # S/fxyz.vhd:120-127: fma(d, inv3, acc), one rounding
                acc[k] = f32(np.float64(d[k]) * np.float64(w) + np.float64(acc[k]))
        assert [float(v) for v in acc] == [float(v) for v in got[i]], i


@settings(**SET)
@given(n=st.integers(1, 100), seed=st.integers(0, 2**31), dt=st.sampled_from([0.01, 0.001, 0.125]))
def test_step_is_kick_then_drift(orc, n, seed, dt):
    # north_star: bodyForce applies v += dt F(x); integrate applies x += dt v with the UPDATED velocity
    b = orc.widen(_bodies(orc, n, seed))
    a = orc.accel_f64(b)
    s1 = orc.run(b, dt, 1)
    k = orc.body_force(b, dt)
    for c, q in zip(("vx", "vy", "vz"), range(3)):
        np.testing.assert_allclose(k[c], b[c] + dt * a[:, q], rtol=1e-14, atol=1e-15 * (1 + np.abs(dt * a[:, q]).max()))   # fma vs mul+add
        assert np.array_equal(k[c.replace("v", "")], b[c.replace("v", "")])       # positions untouched by the kick
    dft = orc.integrate(k, dt)
    for c in "xyz":
        np.testing.assert_allclose(dft[c], k[c] + dt * k["v" + c], rtol=1e-14, atol=1e-15 * (1 + np.abs(dt * k["v" + c]).max()))
    for c in b.dtype.names:
        assert np.array_equal(s1[c], dft[c]), c


def test_single_body_feels_nothing(orc):
    b = orc.widen(_bodies(orc, 1, 3))
    assert np.array_equal(orc.accel_f64(b), np.zeros((1, 3)))
    assert np.array_equal(orc.accel_f32(_bodies(orc, 1, 3)), np.zeros((1, 3), dtype=f32))


def test_fpga_order_and_sequential_order_agree_to_rounding(orc):
    # 16 interleaved partial sums + tree (S/fxyz.vhd:120-145, S/final_adder.vhd:88-104) vs one sequential chain
    b = _bodies(orc, 1000, 9)
    ref = orc.accel_f64_from_f32(b)
    e_seq = orc.rel_err(orc.accel_f32(b), ref).max()
    e_fpga = orc.rel_err(orc.accel_f32(b, order="fpga"), ref).max()
    assert e_seq <= 1e-5 and e_fpga <= 1e-5
