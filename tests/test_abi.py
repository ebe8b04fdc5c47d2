"""Host-side checks that need no GPU: the C-ABI library loads and exports every symbol include/nbody.h
declares, the planner, the error behaviour without a device, and the Body layout."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "nbody.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:int|void|char)\s*\*?\s*(\w+)\s*\(", src, flags=re.M)
    return sorted(set(names))


def test_library_exports_every_declared_symbol(nb):
    names = _declared_functions()
    assert len(names) >= 30
    lib = C.CDLL(nb.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libnbody_b200.so does not export %s" % n
    # and the Python mirror binds exactly that set
    assert sorted(nb.SYMBOLS) == names


def test_body_layout(nb):
    assert nb.body_dtype.itemsize == 24 and nb.bodyd_dtype.itemsize == 48
    assert nb.body_dtype.names == ("x", "y", "z", "vx", "vy", "vz")
    hdr = open(os.path.join(ROOT, "include", "nbody.h")).read()
    assert "typedef struct { float x, y, z, vx, vy, vz; } Body;" in hdr


def test_randomize_matches_oracle_stream(nb, orc):
    a = nb.randomizeBodies(1000, seed=42)
    b = orc.randomize(1000, 42)
    np.testing.assert_array_equal(a.view(np.float32), b.view(np.float32))
    d = nb.randomizeBodies(10, seed=42, dtype=nb.bodyd_dtype)
    np.testing.assert_array_equal(d["vz"], a["vz"][:10].astype(np.float64))
    # the 2-argument reference-shaped form: n counts floats, default seed 42
    raw = np.empty(60, dtype=np.float32)
    nb.lib().randomizeBodies(raw.ctypes.data_as(C.c_void_p), 60)
    np.testing.assert_array_equal(raw, a.view(np.float32)[:60])


@pytest.mark.parametrize("n,world", [(1048576, 1), (1048576, 8), (4194304, 4), (131072, 2), (65536, 1), (4096, 1), (1000, 3), (1, 1), (129, 2)])
def test_plan_partitions_cover_all_bodies(nb, n, world):
    covered = 0
    for r in range(world):
        p = nb.plan(n, rank=r, world=world)
        assert p["blk"] == 128 and p["total_blocks"] == p["local_blocks"] * world
        assert p["total_blocks"] * 128 >= n
        assert (p["i_begin"], p["i_end"]) == nb.shard_range(n, r, world)
        assert p["i_begin"] == covered
        covered = p["i_end"]
        assert 1 <= p["splits_local"] <= 48 and 0 <= p["splits_remote"] <= 48
        assert (p["splits_remote"] == 0) == (world == 1)
        assert p["slots"] == p["splits_local"] + p["splits_remote"] <= 96
        assert p["i_tiles"] * p["tile_bodies"] >= p["local_blocks"] * 128
        # summation chains stay bounded (accuracy): bodies per split <= 65536 + one block
        jl = p["total_blocks"] if world == 1 else p["local_blocks"]
        if p["splits_local"] < 48:
            assert jl * 128 / p["splits_local"] <= 65536 + 128
    assert covered == n


def test_plan_fills_whole_waves(nb):
    # N=1M on 8 GPUs: 128 i-tiles per rank; the planner must cut j so that CTAs come in near-whole waves
    p = nb.plan(1048576, rank=0, world=8, sms=148, variant=3)
    for s, jl in ((p["splits_local"], p["local_blocks"]), (p["splits_remote"], p["total_blocks"] - p["local_blocks"])):
        units = p["i_tiles"] * s
        assert units >= 4 * 148 * 2                      # several waves of CTAs, so the dynamic scheduler can balance
        unit = -(-jl // s)
        assert unit * s <= jl + s                        # even cut: split sizes differ by at most one block


def test_plan_regimes_of_the_split_model(nb):
    # the three regimes of choose_splits (capi.cu), on the shipped 1024-body shape (variant 14, 2 CTAs per SM):
    # few tiles -> one resident wave, preferably one CTA per SM; many long CTAs -> whole waves of 296
    for n in (6144, 8192, 10240, 16384):
        p = nb.plan(n, sms=148, variant=14)
        assert p["i_tiles"] * p["splits_local"] <= 148, (n, p)          # every CTA has an SM to itself
    for n in (12288, 20480, 24576):
        p = nb.plan(n, sms=148, variant=14)
        assert p["i_tiles"] * p["splits_local"] <= 296, (n, p)          # one resident wave
    for n in (65536, 131072, 262144, 1048576):
        p = nb.plan(n, sms=148, variant=14)
        assert (p["i_tiles"] * p["splits_local"]) % 296 == 0, (n, p)     # whole waves
    for world in (2, 4, 8):                                               # both passes of a sharded step as well
        p = nb.plan(1048576, rank=world - 1, world=world, sms=148, variant=14)
        assert (p["i_tiles"] * p["splits_local"]) % 296 == 0 and (p["i_tiles"] * p["splits_remote"]) % 296 == 0, p
    # a machine with a different SM count is planned for its own wave size
    p = nb.plan(131072, sms=132, variant=14)
    assert (p["i_tiles"] * p["splits_local"]) % (2 * 132) == 0, p


@pytest.mark.parametrize("n,world,variant,prec", [(1048576, 1, 19, 0), (1048576, 8, 19, 0), (4194304, 2, 19, 0), (131072, 1, 19, 0),
                                                  (16384, 1, 19, 0), (8192, 1, 19, 0), (10000, 3, 19, 0), (6144, 1, 21, 0),
                                                  (1, 1, 21, 0), (129, 2, 21, 0), (65536, 1, 5, 1), (65536, 8, 5, 1)])
def test_stream_plan_covers_every_pair_once_and_balances(nb, n, world, variant, prec):
    """stream-K decomposition (csrc/stream.cuh), walked on the host with the arithmetic the kernel runs: every
    (tile, phase) j-range is covered exactly once, workspace slots are unique, every CTA gets the same number of
    granule units to within one per phase, and every CTA agrees on how many segments a tile has."""
    for rank in sorted({0, world - 1}):
        p, segs = nb.stream_segments(n, prec, rank, world, 148, variant)
        G, T = p["stream_grid"], p["i_tiles"]
        assert 1 <= G <= 148 * 7 and p["stream_phases"] in ((1,) if world == 1 else (1, 2))
        if world > 1:                                   # own-slice-first only where the step is long against the exchange
            est_us = (p["i_end"] - p["i_begin"] if rank < world - 1 else p["local_blocks"] * 128) * n / (3.1e6 if prec == 0 else 1.08e6)
            assert (p["stream_phases"] == 2) == (min(n, p["local_blocks"] * 128) * n / (3.1e6 if prec == 0 else 1.08e6) >= 1000.0), (p, est_us)
        gl, gt = p["local_blocks"] * 8, p["total_blocks"] * 8
        ph_len = [gt] if p["stream_phases"] == 1 else [gl, gt - gl]
        cover, slots, per_tile, work = {}, set(), {}, {}
        for c, rows in segs.items():
            assert len(rows) <= 4096
            for (ph, t, ja, jb, slot, nseg) in rows:
                assert 0 <= t < T and 0 <= ja < jb <= ph_len[ph]
                cover.setdefault((ph, t), []).append((ja, jb))
                assert slot not in slots and 0 <= slot < p["stream_phases"] * (T + G)
                slots.add(slot)
                per_tile.setdefault(t, set()).add(nseg)
                work[(c, ph)] = work.get((c, ph), 0) + jb - ja
        for ph in range(p["stream_phases"]):
            for t in range(T):
                iv = sorted(cover[(ph, t)])
                assert iv[0][0] == 0 and iv[-1][1] == ph_len[ph]
                assert all(a[1] == b[0] for a, b in zip(iv, iv[1:]))            # no gap, no overlap
            w = [work.get((c, ph), 0) for c in range(G)]
            assert max(w) - min(w) <= 1 and min(w) >= 8                          # >= one layout block of j per CTA and phase
        for t in range(T):
            assert per_tile[t] == {sum(len(cover[(ph, t)]) for ph in range(p["stream_phases"]))}


@pytest.mark.parametrize("i_tiles,nsplit,ring,order", [(1024, 37, 64, 1), (256, 48, 48, 1), (100, 7, 64, 1), (65, 3, 64, 1), (128, 37, 128, 0), (5, 1, 64, 1), (4096, 48, 48, 1)])
def test_fused_pass_cta_map_is_a_bijection_and_reuses_ring_positions_in_order(nb, i_tiles, nsplit, ring, order):
    """fused split-grid pass (csrc/force_f32.cu): CTA -> (tile, split).  Every (tile, split) exactly once; with groups
    (order 1) CTAs run split by split inside a group of ring/2 tiles, groups in order, so that when a CTA of tile t is
    dispatched every CTA of tile t - ring was dispatched at least one whole group earlier (the reuse wait is then a formality);
    the rank's own j-slice (the low splits) comes first inside every group."""
    seen, first_bid, last_bid = set(), {}, {}
    step = 1 if i_tiles * nsplit <= 20000 else 7                     # sample the big grids, walk the small ones
    bids = range(0, i_tiles * nsplit, step)
    prev = None
    for bid in bids:
        t, s = nb.fused_cta(i_tiles, nsplit, ring, order, bid)
        assert 0 <= t < i_tiles and 0 <= s < nsplit and (t, s) not in seen
        seen.add((t, s))
        first_bid.setdefault(t, bid); last_bid[t] = bid
        if order == 1 and prev is not None and step == 1:
            gs = ring // 2
            assert (t // gs, s, t) >= (prev[0] // gs, prev[1], prev[0])          # group, then split, then tile: lexicographic
        prev = (t, s)
    if step == 1:
        assert len(seen) == i_tiles * nsplit
        if order == 1:
            gs = ring // 2
            for t in range(ring, i_tiles):
                assert first_bid[t] - last_bid[t - ring] >= gs * nsplit - gs    # a whole group of CTAs lies in between
    with pytest.raises(nb.NBodyError):
        nb.fused_cta(i_tiles, nsplit, ring, order, i_tiles * nsplit)


def test_plan_rejects_bad_arguments(nb):
    for kw in (dict(n=0), dict(n=16, rank=2, world=2), dict(n=16, precision=7), dict(n=16, variant=999)):
        args = dict(n=16, precision=0, rank=0, world=1, sms=148, variant=0); args.update(kw)
        with pytest.raises(nb.NBodyError):
            nb.plan(**args)


def test_no_gpu_fails_loudly(nb):
    from conftest import HAVE_GPU
    if HAVE_GPU:
        pytest.skip("a GPU is present")
    with pytest.raises(nb.NBodyError, match="no CPU fallback"):
        nb.NBody(1024)


def test_python_mirror_argument_checks(nb):
    with pytest.raises(TypeError):
        nb._as_bodies(np.zeros(4, dtype=np.int32), nb.body_dtype)
    a = nb._as_bodies(np.zeros((5, 6), dtype=np.float32), nb.body_dtype)
    assert a.shape == (5,) and a.dtype == nb.body_dtype
    with pytest.raises(ValueError):
        nb.mailbox_forces(np.zeros((4, 3), dtype=np.float32))
