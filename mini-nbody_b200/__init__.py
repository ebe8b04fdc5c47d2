"""Host-side mirror of the reference interface for mini-nbody's hot path, over the C ABI of
libnbody_b200.so (include/nbody.h).

The reference's host code is C (absent from the mount; contract = BASELINE.json north_star):
``Body{x,y,z,vx,vy,vz}``, ``randomizeBodies``, ``bodyForce``, ``integrate``.  The real host
program is apps/nbody.c; this module is the same surface for Python callers (tests, bench.py):
same names, same argument meaning, same error behaviour (an exception where the C entry point
would abort).  There is no CPU fallback: if libnbody_b200.so is missing or no B200 is visible,
calls fail loudly.
"""
import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NBODY_B200_LIB") or os.path.join(PKG_DIR, "libnbody_b200.so")   # override: A/B builds

F32, F64 = 0, 1
SOFTENING = np.float32(1.0e-9)
NCCL_ID_BYTES = 128
IPC_BLOB_BYTES = 256

# Body{x,y,z,vx,vy,vz}: 24-byte AoS record (48 bytes for the FP64 path)
body_dtype = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("vx", "<f4"), ("vy", "<f4"), ("vz", "<f4")])
bodyd_dtype = np.dtype([("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("vx", "<f8"), ("vy", "<f8"), ("vz", "<f8")])


class NBodyError(RuntimeError):
    pass


class Plan(C.Structure):
    _fields_ = [(k, C.c_int) for k in (
        "n", "world", "rank", "blk", "total_blocks", "local_blocks", "i_begin", "i_end",
        "tile_bodies", "i_tiles", "splits_local", "splits_remote", "slots", "stream_grid", "stream_phases")]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/nbody.h declares: name -> (restype, argtypes)
_vp, _i, _d, _ll = C.c_void_p, C.c_int, C.c_double, C.c_longlong
SYMBOLS = {
    "randomizeBodies": (None, [_vp, _i]),
    "randomizeBodiesSeeded": (None, [_vp, _ll, C.c_uint64]),
    "bodyForce": (None, [_vp, C.c_float, _i]),
    "integrate": (None, [_vp, C.c_float, _i]),
    "bodyForceD": (None, [_vp, _d, _i]),
    "integrateD": (None, [_vp, _d, _i]),
    "nbody_create": (_i, [_i, _i, _i, C.POINTER(_vp)]),
    "nbody_nccl_unique_id": (_i, [_vp]),
    "nbody_create_rank": (_i, [_i, _i, _i, _i, _i, _vp, C.POINTER(_vp)]),
    "nbody_destroy": (_i, [_vp]),
    "nbody_ipc_export": (_i, [_vp, _vp]),
    "nbody_ipc_import": (_i, [_vp, _vp]),
    "nbody_upload": (_i, [_vp, _vp]),
    "nbody_upload_d": (_i, [_vp, _vp]),
    "nbody_download": (_i, [_vp, _vp]),
    "nbody_download_d": (_i, [_vp, _vp]),
    "nbody_download_local": (_i, [_vp, _vp]),
    "nbody_download_local_d": (_i, [_vp, _vp]),
    "nbody_step": (_i, [_vp, _d, _i]),
    "nbody_step_async": (_i, [_vp, _d, _i]),
    "nbody_sync": (_i, [_vp]),
    "nbody_body_force": (_i, [_vp, _d]),
    "nbody_integrate": (_i, [_vp, _d]),
    "nbody_set_softening": (_i, [_vp, _d]),
    "nbody_get_softening": (_i, [_vp, C.POINTER(_d)]),
    "nbody_step_kdk": (_i, [_vp, _d, _i]),
    "nbody_accel": (_i, [_vp, _vp]),
    "nbody_accel_d": (_i, [_vp, _vp]),
    "nbody_energy": (_i, [_vp, C.POINTER(_d), C.POINTER(_d)]),
    "nbody_mailbox_forces": (_i, [_vp, _vp, _i]),
    "nbody_mailbox_run": (_i, [_vp, _vp, _i]),
    "nbody_set_option": (_i, [_vp, C.c_char_p, _ll]),
    "nbody_get_info": (_i, [_vp, C.c_char_p, C.POINTER(_ll)]),
    "nbody_timing_reset": (_i, [_vp]),
    "nbody_timing_get": (_i, [_vp, C.POINTER(_d), C.POINTER(_d), C.POINTER(_ll)]),
    "nbody_last_step_ms": (_i, [_vp, C.POINTER(_d)]),
    "nbody_stream_profile": (_i, [_vp, _vp, _i]),
    "nbody_probe_fp32_peak": (_i, [_vp, C.POINTER(_d), C.POINTER(_d)]),
    "nbody_plan": (_i, [_i, _i, _i, _i, _i, _i, C.POINTER(Plan)]),
    "nbody_stream_segments": (_i, [C.POINTER(Plan), _i, _vp, _i]),
    "nbody_fused_cta": (_i, [_i, _i, _i, _i, _i, C.POINTER(_i), C.POINTER(_i)]),
    "nbody_last_error": (C.c_char_p, []),
    "nbody_version": (C.c_char_p, []),
}

_lib = None


def lib():
    """Load libnbody_b200.so (once).  Raises if it has not been built -- never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NBodyError("%s not found: build it with `python mini-nbody_b200/build.py` "
                             "(there is no CPU fallback)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(l, name)
            f.restype, f.argtypes = res, args
        _lib = l
    return _lib


def _check(rc, what):
    if rc != 0:
        raise NBodyError("%s failed (%d): %s" % (what, rc, lib().nbody_last_error().decode()))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _as_bodies(p, dtype):
    a = np.asarray(p)
    if a.dtype != dtype:
        if a.dtype == dtype[0] and a.ndim == 2 and a.shape[1] == 6:
            a = a.view(dtype).reshape(-1)
        else:
            raise TypeError("expected an array of %s (or an (n,6) array of its scalar type), got %s" % (dtype, a.dtype))
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("body array must be C-contiguous")
    return a.reshape(-1)


# ---- reference-shaped entry points ---------------------------------------------------------------
def randomizeBodies(n_bodies, seed=42, dtype=body_dtype):
    """n_bodies bodies with all 6 components uniform in [-1, 1); FP64 bodies are the same FP32
    stream widened (BASELINE.json configs: seeded random init)."""
    a = np.empty(n_bodies, dtype=body_dtype)
    lib().randomizeBodiesSeeded(_ptr(a), 6 * n_bodies, seed)
    if dtype == bodyd_dtype:
        d = np.empty(n_bodies, dtype=bodyd_dtype)
        for k in body_dtype.names:
            d[k] = a[k]
        return d
    return a


def bodyForce(p, dt, n=None):
    """v += dt * F(x) in place on the host array `p` (Body or BodyD records)."""
    a = np.asarray(p)
    if a.dtype == bodyd_dtype or a.dtype == np.float64:
        a = _as_bodies(p, bodyd_dtype)
        lib().bodyForceD(_ptr(a), float(dt), len(a) if n is None else n)
    else:
        a = _as_bodies(p, body_dtype)
        lib().bodyForce(_ptr(a), float(dt), len(a) if n is None else n)


def integrate(p, dt, n=None):
    """x += dt * v in place on the host array `p`."""
    a = np.asarray(p)
    if a.dtype == bodyd_dtype or a.dtype == np.float64:
        a = _as_bodies(p, bodyd_dtype)
        lib().integrateD(_ptr(a), float(dt), len(a) if n is None else n)
    else:
        a = _as_bodies(p, body_dtype)
        lib().integrate(_ptr(a), float(dt), len(a) if n is None else n)


def plan(n, precision=F32, rank=0, world=1, sms=148, variant=0):
    """Host-only shard/launch plan (no GPU needed)."""
    p = Plan()
    _check(lib().nbody_plan(n, precision, rank, world, sms, variant, C.byref(p)), "nbody_plan")
    return p.as_dict()


def stream_segments(n, precision=F32, rank=0, world=1, sms=148, variant=19):
    """Host-only walk of a stream-K plan: (plan dict, {cta: [(phase, tile, ja, jb, slot, nseg), ...]})."""
    p = Plan()
    _check(lib().nbody_plan(n, precision, rank, world, sms, variant, C.byref(p)), "nbody_plan")
    out = {}
    buf = np.empty((4096, 6), dtype=np.int32)
    for c in range(p.stream_grid):
        k = lib().nbody_stream_segments(C.byref(p), c, _ptr(buf), len(buf))
        if k < 0:
            raise NBodyError("nbody_stream_segments failed: %s" % lib().nbody_last_error().decode())
        out[c] = [tuple(int(x) for x in row) for row in buf[:k]]
    return p.as_dict(), out


def fused_cta(i_tiles, nsplit, ring, order, bid):
    """(tile, split) of CTA bid of the fused split-grid pass -- the kernel's own map, evaluated on the host"""
    t, sp = C.c_int(), C.c_int()
    _check(lib().nbody_fused_cta(i_tiles, nsplit, ring, order, bid, C.byref(t), C.byref(sp)), "nbody_fused_cta")
    return t.value, sp.value


def nccl_unique_id():
    buf = C.create_string_buffer(NCCL_ID_BYTES)
    _check(lib().nbody_nccl_unique_id(buf), "nbody_nccl_unique_id")
    return buf.raw


def mailbox_forces(words):
    """FPGA mailbox image: (n,4) float32 body words {x,y,z,pad} -> (n,4) result words {Fx,Fy,Fz,0}."""
    w = np.ascontiguousarray(words, dtype=np.float32)
    if w.ndim != 2 or w.shape[1] != 4:
        raise ValueError("words must have shape (n, 4)")
    out = np.empty_like(w)
    _check(lib().nbody_mailbox_forces(_ptr(w), _ptr(out), w.shape[0]), "nbody_mailbox_forces")
    return out


MAILBOX_DEPTH = 32768


def mailbox_run(ram, results):
    """The reference's mailbox handshake (S/top_level.vhd:176-272) on its own RAM images: `ram` and `results` are
    (depth, 4) float32/uint32 arrays of 128-bit words, word 0 of `ram` the control word {BEGIN, NUM_PTS}.  Returns 1 while
    BEGIN is 0, else 0 after writing the forces to results[1:N+1] and the completion word to ram[0]."""
    for a in (ram, results):
        if a.ndim != 2 or a.shape[1] != 4 or a.dtype.itemsize != 4 or not a.flags["C_CONTIGUOUS"]:
            raise ValueError("mailbox images are C-contiguous (depth, 4) arrays of 32-bit words")
    if ram.shape != results.shape:
        raise ValueError("ram and results must have the same depth")
    rc = lib().nbody_mailbox_run(_ptr(ram), _ptr(results), ram.shape[0])
    if rc < 0:
        _check(rc, "nbody_mailbox_run")
    return rc


class NBody:
    """Resident-state handle: bodies stay in HBM between steps."""

    def __init__(self, n, precision=F32, ngpus=1, rank=None, world=None, device=None, nccl_id=None):
        self._h = C.c_void_p()
        self.n, self.precision = int(n), int(precision)
        self.dtype = body_dtype if precision == F32 else bodyd_dtype
        self.scalar = np.float32 if precision == F32 else np.float64
        if rank is None:
            _check(lib().nbody_create(self.n, self.precision, int(ngpus), C.byref(self._h)), "nbody_create")
            self.world, self.rank = int(ngpus), 0
        else:
            idbuf = C.create_string_buffer(nccl_id, NCCL_ID_BYTES) if nccl_id is not None else None
            _check(lib().nbody_create_rank(self.n, self.precision, int(rank), int(world), int(device or 0), idbuf,
                                           C.byref(self._h)), "nbody_create_rank")
            self.world, self.rank = int(world), int(rank)

    def close(self):
        if self._h:
            lib().nbody_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def ipc_export(self):
        buf = C.create_string_buffer(IPC_BLOB_BYTES)
        _check(lib().nbody_ipc_export(self._h, buf), "nbody_ipc_export")
        return buf.raw

    def ipc_import(self, blobs):
        """blobs: the ipc_export() results of all ranks, rank-major (list of bytes or one bytes object)."""
        raw = b"".join(blobs) if not isinstance(blobs, (bytes, bytearray)) else bytes(blobs)
        if len(raw) != IPC_BLOB_BYTES * self.world:
            raise ValueError("expected %d blobs of %d bytes" % (self.world, IPC_BLOB_BYTES))
        _check(lib().nbody_ipc_import(self._h, C.create_string_buffer(raw, len(raw))), "nbody_ipc_import")

    def upload(self, p):
        a = _as_bodies(p, self.dtype)
        if len(a) != self.n:
            raise ValueError("expected %d bodies, got %d" % (self.n, len(a)))
        f = lib().nbody_upload if self.precision == F32 else lib().nbody_upload_d
        _check(f(self._h, _ptr(a)), "nbody_upload")

    def download(self, out=None):
        a = np.empty(self.n, dtype=self.dtype) if out is None else _as_bodies(out, self.dtype)
        f = lib().nbody_download if self.precision == F32 else lib().nbody_download_d
        _check(f(self._h, _ptr(a)), "nbody_download")
        return a

    def download_local(self, out=None):
        """this rank's bodies [i_begin, i_end) only (one-GPU handles): no collective, 1/world of the bytes"""
        n_loc = self.info("i_end") - self.info("i_begin")
        a = np.empty(n_loc, dtype=self.dtype) if out is None else out
        if len(a) != n_loc or a.dtype != self.dtype or not a.flags["C_CONTIGUOUS"]:
            raise ValueError("expected a contiguous array of %d bodies" % n_loc)
        f = lib().nbody_download_local if self.precision == F32 else lib().nbody_download_local_d
        _check(f(self._h, _ptr(a)), "nbody_download_local")
        return a

    def step(self, dt, nsteps=1):
        _check(lib().nbody_step(self._h, float(dt), int(nsteps)), "nbody_step")

    def step_async(self, dt, nsteps=1):
        _check(lib().nbody_step_async(self._h, float(dt), int(nsteps)), "nbody_step_async")

    def sync(self):
        _check(lib().nbody_sync(self._h), "nbody_sync")

    def step_kdk(self, dt, nsteps=1):
        _check(lib().nbody_step_kdk(self._h, float(dt), int(nsteps)), "nbody_step_kdk")

    def set_softening(self, eps):
        _check(lib().nbody_set_softening(self._h, float(eps)), "nbody_set_softening")

    def softening(self):
        e = C.c_double()
        _check(lib().nbody_get_softening(self._h, C.byref(e)), "nbody_get_softening")
        return e.value

    def body_force(self, dt):
        _check(lib().nbody_body_force(self._h, float(dt)), "nbody_body_force")

    def integrate(self, dt):
        _check(lib().nbody_integrate(self._h, float(dt)), "nbody_integrate")

    def accel(self):
        a = np.empty((self.n, 3), dtype=self.scalar)
        f = lib().nbody_accel if self.precision == F32 else lib().nbody_accel_d
        _check(f(self._h, _ptr(a)), "nbody_accel")
        return a

    def energy(self):
        ke, pe = C.c_double(), C.c_double()
        _check(lib().nbody_energy(self._h, C.byref(ke), C.byref(pe)), "nbody_energy")
        return ke.value, pe.value

    def set_option(self, key, value):
        _check(lib().nbody_set_option(self._h, key.encode(), int(value)), "nbody_set_option(%s)" % key)

    def info(self, key):
        v = C.c_longlong()
        _check(lib().nbody_get_info(self._h, key.encode(), C.byref(v)), "nbody_get_info(%s)" % key)
        return v.value

    def timing_reset(self):
        _check(lib().nbody_timing_reset(self._h), "nbody_timing_reset")

    def timing(self):
        f, g, l = C.c_double(), C.c_double(), C.c_longlong()
        _check(lib().nbody_timing_get(self._h, C.byref(f), C.byref(g), C.byref(l)), "nbody_timing_get")
        return {"force_ms": f.value, "integrate_ms": g.value, "launches": l.value}

    def stream_profile(self):
        """per-CTA timeline of the last stream-K pass (option profile=1): (G, 8) uint64, see include/nbody.h"""
        rows = np.zeros((65536, 8), dtype=np.uint64)
        g = lib().nbody_stream_profile(self._h, _ptr(rows), len(rows))
        if g < 0:
            raise NBodyError("nbody_stream_profile failed: %s" % lib().nbody_last_error().decode())
        return rows[:g].copy()

    def last_step_ms(self):
        ms = C.c_double()
        _check(lib().nbody_last_step_ms(self._h, C.byref(ms)), "nbody_last_step_ms")
        return ms.value

    def probe_fp32_peak(self):
        a, b = C.c_double(), C.c_double()
        _check(lib().nbody_probe_fp32_peak(self._h, C.byref(a), C.byref(b)), "nbody_probe_fp32_peak")
        return {"ffma_lane_ops_per_s": a.value, "sm_clock_mhz": b.value}


def shard_range(n, rank, world, blk=128):
    """[i_begin, i_end) of the bodies rank `rank` owns -- pure-Python mirror of nbody_plan()."""
    nblocks = (n + blk - 1) // blk
    local = (nblocks + world - 1) // world
    return min(n, rank * local * blk), min(n, (rank + 1) * local * blk)
