"""Sharded path: one process driving G ranks, i-bodies sharded, positions replicated, per-step exchange (NCCL all-gather or
peer-memory push).  With >= 2 visible devices the ranks are one per GPU.  On a one-GPU box the same tests run with
NBODY_VIRTUAL_RANKS=1: all ranks on the one device (slices, push exchange, step flags, in-kernel waits, sharded host I/O all
exercised; only the NCCL transport cannot be, there is no NCCL between ranks of one device) -- so the driver's single-GPU
test tier covers the N > 1 path too instead of skipping it."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
DT = 0.01


@pytest.fixture
def ranks(monkeypatch):
    """(number of ranks, virtual?)"""
    g = min(_ngpu(), 8)
    if g >= 2:
        return g, False
    monkeypatch.setenv("NBODY_VIRTUAL_RANKS", "1")
    return 3, True


def _ngpu():
    try:
        cudart = C.CDLL("libcudart.so")
    except OSError:
        cudart = C.CDLL("/usr/local/cuda/lib64/libcudart.so")
    n = C.c_int(0)
    return n.value if cudart.cudaGetDeviceCount(C.byref(n)) == 0 else 0


@pytest.mark.parametrize("overlap,exchange,fuse", [(1, 0, 0), (0, 0, 0), (1, 1, 0), (0, 1, 0), (1, 1, 1)])
def test_sharded_matches_single_gpu(nb, orc, ranks, overlap, exchange, fuse):
    """fuse = 1: the one-launch pass (rotated j-range, last-arriver reduction + integrate + push in the force kernel, peers'
    flags acquired by the CTAs that leave the rank's own slice); fuse = 0: two force launches around the wait + integrate kernel"""
    g, virtual = ranks
    if virtual and exchange == 0:
        pytest.skip("the NCCL transport needs one GPU per rank")
    n = 40000                                            # not a multiple of 128*g
    b = orc.randomize(n, 42)
    with nb.NBody(n) as h1:
        h1.upload(b); a1 = h1.accel(); h1.step(DT, 1); s1 = h1.download(); e1 = h1.energy()
        h1.step(DT, 2); s1_3 = h1.download()
    with nb.NBody(n, ngpus=g) as hg:
        hg.set_option("overlap", overlap)
        hg.set_option("exchange", exchange)            # 0 = NCCL all-gather, 1 = peer-memory push from the integrate kernel
        hg.set_option("fuse", fuse)
        assert hg.info("exchange") == exchange and hg.info("fuse") == fuse and hg.info("virtual_ranks") == int(virtual)
        two_pass = hg.info("phases") == 2 if hg.info("stream") else hg.info("splits_remote") > 0
        assert hg.info("world") == g and two_pass == (bool(overlap) and not fuse)
        hg.upload(b); ag = hg.accel(); hg.step(DT, 1); sg = hg.download(); eg = hg.energy()
        hg.step(DT, 2); s3 = hg.download()                 # further steps exercise the double-buffered exchange
    assert orc.rel_err(ag, orc.accel_f64_from_f32(b)).max() <= 1e-5
    assert orc.rel_err(ag, a1).max() <= 2e-5       # two FP32 summation orders, each within 1e-5 of the truth
    # one step: v = v0 + dt*a, x = x0 + dt*v -- state agrees to FP32 rounding of |dt*a| ~ 1e2
    for comps in (("x", "y", "z"), ("vx", "vy", "vz")):
        dv = np.sqrt(sum((sg[k].astype(np.float64) - s1[k].astype(np.float64)) ** 2 for k in comps))
        nv = np.sqrt(sum(s1[k].astype(np.float64) ** 2 for k in comps))
        assert (dv / np.maximum(1.0, nv)).max() <= 3e-5, comps
    assert abs(sum(eg) - sum(e1)) <= 1e-4 * abs(sum(e1))
    # three steps: the system is chaotic, so only statistical agreement can be asked for --
    # the sharded run must sit as close to the CPU reference trajectory as the single-GPU run does
    ref3 = orc.run(b, DT, 3)
    assert np.isfinite(s3.view(np.float32)).all()
    def dist3(p):
        return np.sqrt(sum((p[k].astype(np.float64) - ref3[k]) ** 2 for k in "xyz")) / np.maximum(1.0, np.sqrt(sum(ref3[k].astype(np.float64) ** 2 for k in "xyz")))
    assert np.median(dist3(s3)) <= 3.0 * np.median(dist3(s1_3)) + 1e-6


@pytest.mark.parametrize("n,exchange", [(10000, 0), (30000, 0), (30000, 1), (70000, 1)])   # from 8192 bodies per rank: stream-K kernel, two phases
def test_sharded_fp64(nb, orc, ranks, n, exchange):
    g, virtual = ranks
    if virtual and exchange == 0:
        pytest.skip("the NCCL transport needs one GPU per rank")
    b = orc.widen(orc.randomize(n, 1))
    with nb.NBody(n, nb.F64, ngpus=g) as hg:
        hg.set_option("exchange", exchange)
        if n // g >= 8192:
            local = -(-(-(-n // 128)) // g) * 128
            assert hg.info("stream") == 1 and hg.info("phases") == (2 if min(n, local) * n / 1.08e6 >= 1000.0 else 1)   # short steps: one phase
            if n == 30000:
                hg.set_option("overlap", 2)                               # ... two on request (own j-slice first)
                assert hg.info("phases") == 2
        hg.upload(b); a = hg.accel(); hg.step(DT, 2); out = hg.download()
    assert orc.rel_err(a, orc.accel_f64(b)).max() <= 1e-12
    ref = orc.run(b, DT, 2)
    for k in "xyz":
        np.testing.assert_allclose(out[k], ref[k], rtol=1e-9, atol=1e-11)


def test_push_and_allgather_exchange_agree_bitwise(nb, orc):
    g = min(_ngpu(), 8)
    if g < 2:
        pytest.skip("the NCCL transport needs one GPU per rank")
    n = 30000
    b = orc.randomize(n, 17)
    outs = []
    for exchange in (0, 1):
        with nb.NBody(n, ngpus=g) as h:
            h.set_option("exchange", exchange); h.set_option("fuse", 0)      # same kernels and splits under both transports
            h.upload(b); h.step(DT, 4); h.body_force(DT); h.integrate(DT)
            outs.append(h.download().view(np.float32).copy())
    np.testing.assert_array_equal(outs[0], outs[1])    # same kernels, same order: only the transport differs


def test_fused_sharded_pass_is_deterministic_and_matches_the_unfused_one(nb, orc, ranks):
    g, virtual = ranks
    n = 50000
    b = orc.randomize(n, 23)
    outs = []
    for fuse in (1, 1, 0):
        with nb.NBody(n, ngpus=g) as h:
            h.set_option("exchange", 1); h.set_option("fuse", fuse)
            h.upload(b); a0 = h.accel(); h.step(DT, 3)
            outs.append((a0, h.download().view(np.float32).copy(), h.info("launches"), h.accel()))
    np.testing.assert_array_equal(outs[0][1], outs[1][1])          # fixed-order in-kernel reduction: run-to-run identical
    np.testing.assert_array_equal(outs[0][3], outs[1][3])
    assert np.isfinite(outs[0][1]).all()
    assert orc.rel_err(outs[0][0], outs[2][0]).max() <= 4e-6        # same state, two decompositions of the j-sweep
    assert orc.rel_err(outs[0][0][20000:21024], orc.accel_f64_from_f32(b, 20000, 21024)).max() <= 1e-5
    assert outs[0][2] < outs[2][2]                                  # fewer launches: 1 per rank and step instead of 4


def test_one_process_per_gpu_sharded_io(nb, orc):
    """torchrun, one rank per GPU: nbody_upload reads only the rank's slice of the host array (the other slices are
    NaN-poisoned on every rank), nbody_download_local returns the rank's slice of the full nbody_download."""
    import os, subprocess, sys
    g = min(_ngpu(), 4)
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(g), "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(root, "tests", "mp_local_io.py")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "MP_LOCAL_IO_OK" in r.stdout, r.stdout[-3000:]


def test_download_local_on_a_single_gpu_handle(nb, orc):
    n = 5000
    b = orc.randomize(n, 3)
    with nb.NBody(n) as h:
        h.upload(b); h.step(DT, 2)
        full, mine = h.download(), h.download_local()
        assert (h.info("i_begin"), h.info("i_end")) == (0, n)
    assert all(np.array_equal(full[k], mine[k]) for k in full.dtype.names)
    if _ngpu() >= 2:
        with nb.NBody(n, ngpus=2) as h:
            h.upload(b)
            with pytest.raises(nb.NBodyError):
                h.download_local()                              # a handle that drives several GPUs has no single slice
