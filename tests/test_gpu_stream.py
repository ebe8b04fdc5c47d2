"""The stream-K force pass (csrc/stream.cuh): one persistent launch per force pass, every CTA an equal share of
the (i-tile, j-granule) space, tiles cut by CTA boundaries reduced by the last-arriving CTA in fixed slot order,
integrate fused behind the reduction.  Parity against the oracle through the C ABI, and bit-identity of the
in-kernel reduction with its two-launch twin (every segment to the workspace, a separate kernel adding them in
the same order) -- the reduction the reference does on chip (S/final_adder.vhd:88-104)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
DT = 0.01


def _want_stream(h, prec=0):
    """stream-K is the default for FP64 from 8192 bodies per GPU and an option for FP32 (variants 19-24 / 5-7)"""
    h.set_option("stream", 1)
    if not h.info("stream"):
        h.set_option("variant", 5 if prec else 19)
    assert h.info("stream") == 1


def _state_and_accel(nb, b, prec, steps=3, **opts):
    with nb.NBody(len(b), prec) as h:
        _want_stream(h, prec)
        for k, v in opts.items():
            h.set_option(k, v)
        h.upload(b)
        a = h.accel()
        h.step(DT, steps)
        h.body_force(DT)
        out = h.download()
        info = {k: h.info(k) for k in ("stream", "grid", "phases", "variant", "i_tiles")}
    return a, out, info


@pytest.mark.parametrize("n,prec,grid", [(6144, 0, 0), (10000, 0, 0), (10000, 0, 7), (33000, 0, 100), (33000, 0, 296), (131072, 0, 0),
                                         (5000, 1, 0), (10000, 1, 37), (40000, 1, 0)])
def test_in_kernel_reduction_is_bit_identical_to_its_two_launch_twin(nb, orc, n, prec, grid):
    b = orc.randomize(n, 77 + n)
    if prec:
        b = orc.widen(b)
    a1, s1, i1 = _state_and_accel(nb, b, prec, grid=grid)
    a2, s2, i2 = _state_and_accel(nb, b, prec, grid=grid, stream_twin=1)
    assert i1["stream"] == 1 and i1 == i2
    if grid:
        assert i1["grid"] == grid
    assert np.array_equal(a1, a2)
    for k in s1.dtype.names:
        assert np.array_equal(s1[k], s2[k]), k
    ref = orc.accel_f64(b) if prec else orc.accel_f64_from_f32(b)
    assert orc.rel_err(a1, ref).max() <= (1e-12 if prec else 1e-5)


@pytest.mark.parametrize("n,prec,grid", [(10000, 0, 0), (33000, 0, 100), (131072, 0, 0), (10000, 1, 37), (40000, 1, 0)])
def test_cooperative_and_last_arriver_reductions_agree_bitwise(nb, orc, n, prec, grid):
    """cut tiles are reduced by the tile's last arriver alone (coop = 0, default) or by their contributors, each its share of
    the bodies (coop = 1: every CTA waits for the tile's other segments, safe because all CTAs of the pass are resident): the
    segments are added in slot order either way"""
    b = orc.randomize(n, 99 + n)
    if prec:
        b = orc.widen(b)
    a1, s1, _ = _state_and_accel(nb, b, prec, grid=grid, coop=1)
    a0, s0, _ = _state_and_accel(nb, b, prec, grid=grid, coop=0)
    assert np.array_equal(a1, a0)
    for k in s1.dtype.names:
        assert np.array_equal(s1[k], s0[k]), k


def test_stream_pass_is_deterministic_and_grid_independent_within_tolerance(nb, orc):
    n = 50000
    b = orc.randomize(n, 3)
    ref = orc.accel_f64_from_f32(b, 0, 4096)
    runs = {}
    for grid in (0, 0, 1, 148, 293):
        with nb.NBody(n) as h:
            _want_stream(h); h.set_option("grid", grid); h.upload(b); a = h.accel()
        if grid in runs:
            np.testing.assert_array_equal(a, runs[grid])            # fixed-order reduction: run-to-run identical
        runs[grid] = a
        assert orc.rel_err(a[:4096], ref).max() <= 1e-5, grid
    assert orc.rel_err(runs[1], runs[293]).max() <= 4e-6              # different segment cuts, both within tolerance


def test_long_segments_use_the_third_accumulation_level(nb, orc):
    """grid = 2 at N = 300 000: each CTA sweeps ~150 000 j per tile in ONE segment (> 65 536, the chain bound the
    split-grid kernel enforced with its j-splits); the close-pair body must stay within tolerance."""
    n = 300000
    b = orc.randomize(n, 11)
    b["x"][1000], b["y"][1000], b["z"][1000] = b["x"][7] + 3e-4, b["y"][7], b["z"][7]      # a 3e-4 close pair
    with nb.NBody(n) as h:
        _want_stream(h); h.set_option("grid", 2); h.upload(b); a = h.accel()
        assert h.info("grid") == 2
    for lo in (0, 896):
        ref = orc.accel_f64_from_f32(b, lo, lo + 256)
        assert orc.rel_err(a[lo:lo + 256], ref).max() <= 1e-5


@pytest.mark.parametrize("n", [6144, 6145, 8191, 12345, 20001, 32767])
def test_stream_ragged_sizes_one_step_state(nb, orc, n):
    b = orc.randomize(n, n)
    with nb.NBody(n) as h:
        _want_stream(h)
        h.upload(b); a = h.accel(); h.step(DT, 1); out = h.download()
    ref64 = orc.accel_f64_from_f32(b)
    assert orc.rel_err(a, ref64).max() <= 1e-5
    # the fused epilogue is the oracle's integrate: v = fma(dt, a, v); x = fma(v, dt, x) on the GPU's own accelerations
    for i, k in enumerate("xyz"):
        vv = (b["v" + k].astype(np.float64) + np.float64(np.float32(DT)) * a[:, i].astype(np.float64)).astype(np.float32)
        assert np.array_equal(out["v" + k], vv), k
        xx = (b[k].astype(np.float64) + vv.astype(np.float64) * np.float64(np.float32(DT))).astype(np.float32)
        assert np.array_equal(out[k], xx), k


def test_stream_and_split_grid_paths_agree_within_tolerance(nb, orc):
    n = 20000
    b = orc.randomize(n, 5)
    outs = {}
    for stream in (1, 0):
        with nb.NBody(n) as h:
            h.set_option("stream", stream); h.set_option("fuse", 0); h.upload(b)
            assert h.info("stream") == stream
            a = h.accel(); h.step(DT, 1); outs[stream] = (a, h.download(), h.info("launches"))
    assert orc.rel_err(outs[1][0], outs[0][0]).max() <= 4e-6
    amax = np.abs(outs[0][0]).max()
    for k in "xyz":
        assert np.abs(outs[1][1][k] - outs[0][1][k]).max() <= 4e-6 * DT * DT * amax + 1e-6      # x += dt*(v + dt*a)


# ---- fused mode of the (i-tile, j-split) grid kernels: last-arriver reduction + integrate in the force kernel --------------
@pytest.mark.parametrize("n,splits,variant,order", [(6144, 0, -1, 1), (10000, 7, -1, 0), (33000, 0, -1, 1), (131072, 0, -1, -1), (20000, 5, 1, 1), (3000, 0, 4, 0),
                                                     (3000, 3, 6, 1), (50000, 48, 13, 1)])
def test_fused_split_grid_pass_is_bit_identical_to_slot_array_plus_integrate_kernel(nb, orc, n, splits, variant, order):
    """Same kernel instantiation, same j-splits, the tile's slots added in the same (split) order: the CTA that completes
    a tile must produce exactly what integrate_kernel produces from the slot array -- accelerations, velocities and
    positions, over several steps (ring positions get reused from the second pass on)."""
    b = orc.randomize(n, 5 + n)
    out = {}
    for fuse in (1, 0):
        with nb.NBody(n) as h:
            h.set_option("stream", 0); h.set_option("fuse", fuse); h.set_option("order", order); h.set_option("graph", 0); h.set_option("fused", 0)
            if variant >= 0:
                h.set_option("variant", variant)
            if splits:
                h.set_option("splits", splits)
            h.upload(b)
            assert h.info("fuse") == fuse and h.info("stream") == 0
            a = h.accel(); h.step(DT, 3); h.body_force(DT); h.step(DT, 2)
            out[fuse] = (a, h.download(), h.info("splits_local"), h.info("launches"))
    assert out[0][2] == out[1][2]
    assert np.array_equal(out[0][0], out[1][0])
    for k in out[0][1].dtype.names:
        assert np.array_equal(out[0][1][k], out[1][1][k]), k
    assert out[1][3] < out[0][3]                               # one launch per pass instead of two
    assert orc.rel_err(out[1][0], orc.accel_f64_from_f32(b)).max() <= 1e-5


def test_fused_pass_small_ring_forces_slot_reuse(nb, orc):
    """N = 262144 with 48 splits: 256 tiles share a ring of 64 positions, every position is reused 4 times per pass."""
    n = 262144
    b = orc.randomize(n, 9)
    with nb.NBody(n) as h:
        h.set_option("stream", 0); h.set_option("splits", 48); h.set_option("order", 1); h.upload(b)
        assert h.info("fuse") == 1 and h.info("order") == 1 and h.info("ring") < h.info("i_tiles")
        a1 = h.accel(); a2 = h.accel()
    np.testing.assert_array_equal(a1, a2)
    ref = orc.accel_f64_from_f32(b, 130000, 130512)
    assert orc.rel_err(a1[130000:130512], ref).max() <= 1e-5
