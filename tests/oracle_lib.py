"""ctypes binding of the CPU oracle (oracle/nbody_oracle.c).  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; the product
package never imports it."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "oracle", "_build")

body_dtype = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("vx", "<f4"), ("vy", "<f4"), ("vz", "<f4")])
bodyd_dtype = np.dtype([("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("vx", "<f8"), ("vy", "<f8"), ("vz", "<f8")])

_libs = {}


def _build():
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("building the oracle failed")


def load(flavour="parity"):
    if flavour not in _libs:
        so = os.path.join(BUILD, "liboracle_%s.so" % flavour)
        if not os.path.exists(so):
            _build()
        l = C.CDLL(so)
        vp, i, f, d, ll = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_longlong
        sig = {
            "oracle_num_threads": (i, []),
            "oracle_set_num_threads": (None, [i]),
            "oracle_accel_f32_kahan": (None, [vp, i, i, i, vp]),
            "oracle_softening_bits": (C.c_uint32, []),
            "oracle_set_softening": (None, [d]),
            "oracle_get_softening": (d, []),
            "oracle_randomize": (None, [vp, ll, C.c_uint64]),
            "oracle_dxy": (f, [f, f, f, f, vp, vp]),
            "oracle_dzsoft": (f, [f, f, vp]),
            "oracle_dxyz_soft": (f, [vp, vp, vp]),
            "oracle_rsqrt": (f, [f]),
            "oracle_cube": (f, [f]),
            "oracle_accel_f32": (None, [vp, i, i, i, vp]),
            "oracle_accel_f32_fpga_order": (None, [vp, i, i, i, vp]),
            "oracle_accel_f64_from_f32": (None, [vp, i, i, i, vp]),
            "oracle_accel_f64": (None, [vp, i, i, i, vp]),
            "oracle_accel_f80": (None, [vp, i, i, i, vp]),
            "oracle_body_force_f32": (None, [vp, f, i]),
            "oracle_integrate_f32": (None, [vp, f, i]),
            "oracle_body_force_f64": (None, [vp, d, i]),
            "oracle_integrate_f64": (None, [vp, d, i]),
            "oracle_run_f32": (None, [vp, f, i, i]),
            "oracle_run_f64": (None, [vp, d, i, i]),
            "oracle_energy_f64": (None, [vp, i, C.POINTER(d), C.POINTER(d)]),
            "oracle_energy_f32in": (None, [vp, i, C.POINTER(d), C.POINTER(d)]),
            "oracle_body_force_f32_fast": (d, [vp, f, i, i, i]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _libs[flavour] = l
    return _libs[flavour]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class softening:
    """with oracle_lib.softening(1e-2): ... -- force loops and energy of the oracle use this softening inside"""
    def __init__(self, eps):
        self.eps = eps

    def __enter__(self):
        self.old = load().oracle_get_softening(); load().oracle_set_softening(self.eps)
        return self

    def __exit__(self, *a):
        load().oracle_set_softening(self.old)


def randomize(n_bodies, seed=42):
    a = np.empty(n_bodies, dtype=body_dtype)
    load().oracle_randomize(_p(a), 6 * n_bodies, seed)
    return a


def widen(b):
    d = np.empty(len(b), dtype=bodyd_dtype)
    for k in body_dtype.names:
        d[k] = b[k]
    return d


def accel_f32(b, i0=0, i1=None, order="sequential"):
    i1 = len(b) if i1 is None else i1
    out = np.empty((i1 - i0, 3), dtype=np.float32)
    fn = {"sequential": load().oracle_accel_f32, "fpga": load().oracle_accel_f32_fpga_order, "kahan": load().oracle_accel_f32_kahan}[order]
    fn(_p(b), len(b), i0, i1, _p(out))
    return out


def accel_f64_from_f32(b, i0=0, i1=None):
    i1 = len(b) if i1 is None else i1
    out = np.empty((i1 - i0, 3), dtype=np.float64)
    load().oracle_accel_f64_from_f32(_p(b), len(b), i0, i1, _p(out))
    return out


def accel_f64(b, i0=0, i1=None, extended=False):
    i1 = len(b) if i1 is None else i1
    out = np.empty((i1 - i0, 3), dtype=np.float64)
    (load().oracle_accel_f80 if extended else load().oracle_accel_f64)(_p(b), len(b), i0, i1, _p(out))
    return out


def run(b, dt, steps):
    b = b.copy()
    if b.dtype == body_dtype:
        load().oracle_run_f32(_p(b), dt, len(b), steps)
    else:
        load().oracle_run_f64(_p(b), dt, len(b), steps)
    return b


def body_force(b, dt):
    b = b.copy()
    (load().oracle_body_force_f32 if b.dtype == body_dtype else load().oracle_body_force_f64)(_p(b), dt, len(b))
    return b


def integrate(b, dt):
    b = b.copy()
    (load().oracle_integrate_f32 if b.dtype == body_dtype else load().oracle_integrate_f64)(_p(b), dt, len(b))
    return b


def energy(b):
    ke, pe = C.c_double(), C.c_double()
    if b.dtype == body_dtype:
        load().oracle_energy_f32in(_p(b), len(b), C.byref(ke), C.byref(pe))
    else:
        load().oracle_energy_f64(_p(b), len(b), C.byref(ke), C.byref(pe))
    return ke.value, pe.value


def rel_err(a, ref):
    """per-body ||a - ref||_2 / ||ref||_2"""
    a = np.asarray(a, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    return np.linalg.norm(a - ref, axis=1) / np.linalg.norm(ref, axis=1)
