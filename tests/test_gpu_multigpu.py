"""Sharded path on real GPUs (needs >= 2 visible devices; skipped otherwise): one process driving G
devices, i-bodies sharded, positions replicated, NCCL all-gather per step."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
DT = 0.01


def _ngpu():
    try:
        cudart = C.CDLL("libcudart.so")
    except OSError:
        cudart = C.CDLL("/usr/local/cuda/lib64/libcudart.so")
    n = C.c_int(0)
    return n.value if cudart.cudaGetDeviceCount(C.byref(n)) == 0 else 0


@pytest.mark.parametrize("overlap", [1, 0])
def test_sharded_matches_single_gpu(nb, orc, overlap):
    g = min(_ngpu(), 8)
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 40000                                            # not a multiple of 128*g
    b = orc.randomize(n, 42)
    with nb.NBody(n) as h1:
        h1.upload(b); a1 = h1.accel(); h1.step(DT, 3); s1 = h1.download(); e1 = h1.energy()
    with nb.NBody(n, ngpus=g) as hg:
        hg.set_option("overlap", overlap)
        hg.upload(b); ag = hg.accel(); hg.step(DT, 3); sg = hg.download(); eg = hg.energy()
    assert orc.rel_err(ag, orc.accel_f64_from_f32(b)).max() <= 1e-5
    assert orc.rel_err(ag, a1).max() <= 2e-5       # two FP32 summation orders, each within 1e-5 of the truth
    d = np.abs(sg.view(np.float32).astype(np.float64) - s1.view(np.float32).astype(np.float64))
    assert np.median(d) <= 1e-6
    assert abs(sum(eg) - sum(e1)) <= 1e-6 * abs(sum(e1))


def test_sharded_fp64(nb, orc):
    g = min(_ngpu(), 8)
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 10000
    b = orc.widen(orc.randomize(n, 1))
    with nb.NBody(n, nb.F64, ngpus=g) as hg:
        hg.upload(b); a = hg.accel(); hg.step(DT, 2); out = hg.download()
    assert orc.rel_err(a, orc.accel_f64(b)).max() <= 1e-12
    ref = orc.run(b, DT, 2)
    for k in "xyz":
        np.testing.assert_allclose(out[k], ref[k], rtol=1e-9, atol=1e-11)
