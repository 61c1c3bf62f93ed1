"""ctypes binding of libmmsig.so (include/mmsig.h).  No torch types cross this boundary:
plain pointers and sizes only.  There is no CPU fallback: a missing library or device raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MMSIG_LIB: developer override used for A/B timing of build variants (profiles/ab_variants.sh)
LIB_PATH = os.environ.get("MMSIG_LIB") or os.path.join(_HERE, "libmmsig.so")

c_dp = C.POINTER(C.c_double)
c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_u8p = C.POINTER(C.c_uint8)
c_u32p = C.POINTER(C.c_uint32)

FLAG_UPDATE_SIGMA, FLAG_FREEZE_TOPICS, FLAG_FREEZE_MU, FLAG_UNSMOOTHED, FLAG_AUTO_ALPHA = 1, 2, 4, 8, 16
STOP_NLOPT27, STOP_NLOPT26 = 0, 1
PRECISION_FP64, PRECISION_FP32 = 0, 1          # mmsig_config.precision (include/mmsig.h)


def _precision(p):
    """0 / 1, or "fp64" / "fp32"."""
    if isinstance(p, str):
        return {"fp64": PRECISION_FP64, "f64": PRECISION_FP64, "fp32": PRECISION_FP32, "f32": PRECISION_FP32}[p.lower()]
    return int(p)
DENSE_TERM_MAJOR, DENSE_SAMPLE_MAJOR = 0, 1


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("stop_rule", C.c_int32), ("profile", C.c_int32), ("precision", C.c_int32),
                ("reserved", C.c_int32 * 4)]


class MmsigError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libmmsig error %d: %s" % (code, msg))
        self.code = code


_SIGS = {
    "mmsig_version": (C.c_int32, []),
    "mmsig_limits": (C.c_int32, [c_i32p] * 5),
    "mmsig_create": (C.c_int32, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "mmsig_destroy": (C.c_int32, [C.c_void_p]),
    "mmsig_last_error": (C.c_char_p, [C.c_void_p]),
    "mmsig_set_stream": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "mmsig_synchronize": (C.c_int32, [C.c_void_p]),
    "mmsig_set_profile": (C.c_int32, [C.c_void_p, C.c_int32]),
    "mmsig_host_alloc": (C.c_int32, [C.c_uint64, C.POINTER(C.c_void_p)]),
    "mmsig_host_free": (C.c_int32, [C.c_void_p]),
    "mmsig_comm_unique_id": (C.c_int32, [c_u8p]),
    "mmsig_comm_init": (C.c_int32, [C.c_void_p, c_u8p, C.c_int32, C.c_int32]),
    "mmsig_group_create": (C.c_int32, [C.POINTER(Config), C.c_int32, c_i32p, C.POINTER(C.c_void_p)]),
    "mmsig_group_destroy": (C.c_int32, [C.c_void_p]),
    "mmsig_group_last_error": (C.c_char_p, [C.c_void_p]),
    "mmsig_group_size": (C.c_int32, [C.c_void_p]),
    "mmsig_group_member": (C.c_void_p, [C.c_void_p, C.c_int32]),
    "mmsig_group_mmctm_set_data": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int32, c_i32p, c_i32p, C.POINTER(c_i64p), C.POINTER(c_i32p), C.POINTER(c_i32p)]),
    "mmsig_group_mmctm_set_state": (C.c_int32, [C.c_void_p] + [c_dp] * 7),
    "mmsig_group_mmctm_iterate": (C.c_int32, [C.c_void_p, C.c_uint32, c_dp]),
    "mmsig_group_mmctm_fit": (C.c_int32, [C.c_void_p, C.c_int32, C.c_double, C.c_uint32, c_dp, c_i32p, c_i32p]),
    "mmsig_group_mmctm_elbo": (C.c_int32, [C.c_void_p, c_dp, c_dp]),
    "mmsig_group_mmctm_get_state": (C.c_int32, [C.c_void_p] + [c_dp] * 10),
    "mmsig_group_mmctm_get_evals": (C.c_int32, [C.c_void_p, c_i32p, c_i32p]),
    "mmsig_group_mmctm_get_theta": (C.c_int32, [C.c_void_p, C.c_int32, c_dp]),
    "mmsig_group_mmctm_fit_host": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int32, c_i32p, c_i32p, C.POINTER(c_i64p), C.POINTER(c_i32p), C.POINTER(c_i32p)] + [c_dp] * 7 +
                                   [C.c_int32, C.c_double, C.c_uint32, c_dp, c_i32p, c_i32p] + [c_dp] * 10),
    "mmsig_group_mmctm_fit_host_packed": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int32, c_i32p, c_i32p, C.POINTER(c_i64p), C.POINTER(c_u32p)] +
                                          [c_dp] * 7 + [C.c_int32, C.c_double, C.c_uint32, c_dp, c_i32p, c_i32p] + [c_dp] * 10),
    "mmsig_group_mmctm_restarts": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int32, c_i32p, c_i32p, C.POINTER(c_i64p), C.POINTER(c_i32p), C.POINTER(c_i32p), c_dp, C.c_int32, c_dp,
                                               C.c_int32, C.c_double, C.c_uint32, c_dp, c_dp, c_i32p, c_i32p]),
    "mmsig_group_lda_set_data": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, c_i64p, c_i32p, c_i32p]),
    "mmsig_group_lda_set_state": (C.c_int32, [C.c_void_p, C.c_double, C.c_double, c_dp, c_dp]),
    "mmsig_group_lda_fit": (C.c_int32, [C.c_void_p, C.c_int32, C.c_double, c_dp, c_i32p, c_i32p]),
    "mmsig_group_lda_elbo": (C.c_int32, [C.c_void_p, c_dp, c_dp]),
    "mmsig_group_lda_get_state": (C.c_int32, [C.c_void_p] + [c_dp] * 6),
    "mmsig_tsv_dims": (C.c_int32, [C.c_char_p, c_i64p, c_i64p]),
    "mmsig_tsv_read": (C.c_int32, [C.c_char_p, C.c_int64, C.c_int64, c_i32p]),
    "mmsig_format_counts": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, c_i64p, c_i64p]),
    "mmsig_format_counts_fetch": (C.c_int32, [C.c_void_p, c_i32p, c_i32p]),
    "mmsig_mmctm_set_data_dense": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, c_i32p, c_i32p,
                                               C.POINTER(C.c_void_p), C.c_int32, C.c_int32]),
    "mmsig_lda_set_data_dense": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                             C.c_int32, C.c_int32]),
    "mmsig_mmctm_set_data": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, c_i32p, c_i32p,
                                         C.POINTER(c_i64p), C.POINTER(c_i32p), C.POINTER(c_i32p)]),
    "mmsig_mmctm_set_state": (C.c_int32, [C.c_void_p] + [c_dp] * 7),
    "mmsig_mmctm_get_alpha": (C.c_int32, [C.c_void_p, c_dp]),
    "mmsig_mmctm_set_phi": (C.c_int32, [C.c_void_p, c_dp]),
    "mmsig_mmctm_iterate": (C.c_int32, [C.c_void_p, C.c_uint32, c_dp]),
    "mmsig_mmctm_fit": (C.c_int32, [C.c_void_p, C.c_int32, C.c_double, C.c_uint32, c_dp, c_i32p, c_i32p]),
    "mmsig_mmctm_fit_host": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, c_i32p, c_i32p,
                                         C.POINTER(c_i64p), C.POINTER(c_i32p), C.POINTER(c_i32p)] + [c_dp] * 7 +
                             [C.c_int32, C.c_double, C.c_uint32, c_dp, c_i32p, c_i32p] + [c_dp] * 10),
    "mmsig_mmctm_fit_host_packed": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, c_i32p, c_i32p,
                                                C.POINTER(c_i64p), C.POINTER(c_u32p)] + [c_dp] * 7 +
                                    [C.c_int32, C.c_double, C.c_uint32, c_dp, c_i32p, c_i32p] + [c_dp] * 10),
    "mmsig_pack_records": (C.c_int32, [C.c_int64, c_i32p, c_i32p, c_u32p]),
    "mmsig_mmctm_elbo": (C.c_int32, [C.c_void_p, c_dp, c_dp]),
    "mmsig_mmctm_get_state": (C.c_int32, [C.c_void_p] + [c_dp] * 10),
    "mmsig_mmctm_get_theta": (C.c_int32, [C.c_void_p, C.c_int32, c_dp]),
    "mmsig_mmctm_restarts": (C.c_int32, [C.c_void_p, C.c_int32, c_dp, C.c_int32, C.c_double, C.c_uint32, c_dp, c_dp,
                                         c_i32p, c_i32p]),
    "mmsig_mmctm_get_evals": (C.c_int32, [C.c_void_p, c_i32p, c_i32p]),
    "mmsig_immctm_set_features": (C.c_int32, [C.c_void_p, c_i32p, C.POINTER(c_i32p)]),
    "mmsig_immctm_set_state": (C.c_int32, [C.c_void_p] + [c_dp] * 7),
    "mmsig_immctm_get_tables": (C.c_int32, [C.c_void_p, c_dp, c_dp, c_dp]),
    "mmsig_lda_set_data": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, c_i64p, c_i32p, c_i32p]),
    "mmsig_lda_set_state": (C.c_int32, [C.c_void_p, C.c_double, C.c_double, c_dp, c_dp]),
    "mmsig_ilda_set_features": (C.c_int32, [C.c_void_p, C.c_int32, c_i32p]),
    "mmsig_ilda_set_state": (C.c_int32, [C.c_void_p, C.c_double, c_dp, c_dp, c_dp]),
    "mmsig_ilda_get_tables": (C.c_int32, [C.c_void_p, c_dp, c_dp]),
    "mmsig_lda_iterate": (C.c_int32, [C.c_void_p, c_dp]),
    "mmsig_lda_set_beta": (C.c_int32, [C.c_void_p, c_dp]),
    "mmsig_lda_iterate_flags": (C.c_int32, [C.c_void_p, C.c_uint32, c_dp]),
    "mmsig_lda_fit": (C.c_int32, [C.c_void_p, C.c_int32, C.c_double, c_dp, c_i32p, c_i32p]),
    "mmsig_lda_fit_host": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, c_i64p, c_i32p, c_i32p, C.c_double, C.c_double,
                                       c_dp, c_dp, C.c_int32, C.c_double, c_dp, c_i32p, c_i32p] + [c_dp] * 6),
    "mmsig_lda_elbo": (C.c_int32, [C.c_void_p, c_dp, c_dp]),
    "mmsig_lda_get_state": (C.c_int32, [C.c_void_p] + [c_dp] * 6),
    "mmsig_lda_get_phi": (C.c_int32, [C.c_void_p, c_dp]),
    "mmsig_debug_math": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int64, c_dp, c_dp]),
    "mmsig_launch_count": (C.c_int64, [C.c_void_p]),
    "mmsig_kernel_times": (C.c_int32, [C.c_void_p, C.c_int32, C.POINTER(C.c_char_p), c_dp, c_i64p, C.c_int32]),
}
EXPORTS = sorted(_SIGS)
_LIB = None


def build_if_stale():
    """(Re)build libmmsig.so with nvcc when it is missing or older than its sources.  Building is
    not computing: the library still refuses to run without a CUDA device."""
    import shutil
    import subprocess
    src_dir = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith((".cu", ".cuh", ".inl"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "mmsig.h"))
    if os.environ.get("MMSIG_LIB"):
        return
    stale = not os.path.exists(LIB_PATH) or any(os.path.getmtime(f) > os.path.getmtime(LIB_PATH) for f in srcs)
    if stale and (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        subprocess.check_call(["make", "-s", "-C", src_dir])


def load():
    """dlopen libmmsig.so and declare every entry point of include/mmsig.h."""
    global _LIB
    if _LIB is None:
        build_if_stale()
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C multimodalmusig.jl_b200/csrc); there is no CPU fallback" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            f = getattr(lib, name)
            f.restype = res
            f.argtypes = args
        _LIB = lib
    return _LIB


def dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def f64(a, n=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None and a.size != n:
        raise ValueError("expected %d values, got %d" % (n, a.size))
    return a


class Handle:
    """Owns one mmsig_handle (one GPU)."""

    def __init__(self, device=0, stop_rule=STOP_NLOPT27, profile=False, precision=PRECISION_FP64):
        self.lib = load()
        cfg = Config(device=device, stop_rule=stop_rule, profile=int(profile), precision=_precision(precision))
        hp = C.c_void_p()
        rc = self.lib.mmsig_create(C.byref(cfg), C.byref(hp))
        if rc != 0:
            raise MmsigError(rc, (self.lib.mmsig_last_error(None) or b"").decode())
        self.h = hp

    def check(self, rc):
        if rc != 0:
            raise MmsigError(rc, (self.lib.mmsig_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.mmsig_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        self.check(self.lib.mmsig_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        self.check(self.lib.mmsig_synchronize(self.h))

    def set_profile(self, on):
        self.check(self.lib.mmsig_set_profile(self.h, int(bool(on))))

    def comm_init(self, uid_bytes, rank, nranks):
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(uid_bytes))
        self.check(self.lib.mmsig_comm_init(self.h, buf, rank, nranks))

    def launch_count(self):
        return int(self.lib.mmsig_launch_count(self.h))

    def kernel_times(self, reset=False):
        n = 64
        names = (C.c_char_p * n)()
        ms = np.zeros(n)
        cnt = np.zeros(n, dtype=np.int64)
        k = self.lib.mmsig_kernel_times(self.h, n, names, dp(ms), cnt.ctypes.data_as(c_i64p), int(reset))
        if k < 0:
            self.check(k)
        return {names[i].decode(): (float(ms[i]), int(cnt[i])) for i in range(k)}


class Group:
    """Owns one mmsig_group: several GPUs driven from this process (include/mmsig.h)."""

    def __init__(self, devices, stop_rule=STOP_NLOPT27, profile=False, precision=PRECISION_FP64):
        self.lib = load()
        self.devices = [int(d) for d in devices]
        cfg = Config(device=self.devices[0], stop_rule=stop_rule, profile=int(profile), precision=_precision(precision))
        ids = np.asarray(self.devices, np.int32)
        gp = C.c_void_p()
        rc = self.lib.mmsig_group_create(C.byref(cfg), len(self.devices), ids.ctypes.data_as(c_i32p), C.byref(gp))
        if rc != 0:
            raise MmsigError(rc, (self.lib.mmsig_group_last_error(None) or b"").decode())
        self.g = gp

    def check(self, rc):
        if rc != 0:
            raise MmsigError(rc, (self.lib.mmsig_group_last_error(self.g) or b"").decode())

    def close(self):
        if getattr(self, "g", None):
            self.lib.mmsig_group_destroy(self.g)
            self.g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def launch_count(self):
        return sum(int(self.lib.mmsig_launch_count(C.c_void_p(self.lib.mmsig_group_member(self.g, i))))
                   for i in range(len(self.devices)))

    def kernel_times(self, member=0, reset=False):
        n = 64
        names = (C.c_char_p * n)()
        ms = np.zeros(n)
        cnt = np.zeros(n, dtype=np.int64)
        h = C.c_void_p(self.lib.mmsig_group_member(self.g, member))
        k = self.lib.mmsig_kernel_times(h, n, names, dp(ms), cnt.ctypes.data_as(c_i64p), int(reset))
        return {names[i].decode(): (float(ms[i]), int(cnt[i])) for i in range(max(k, 0))}


def host_array(shape, dtype):
    """numpy array in page-locked host memory owned by the library (mmsig_host_alloc); freed with the array."""
    lib = load()
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    p = C.c_void_p()
    rc = lib.mmsig_host_alloc(max(n, 1) * dtype.itemsize, C.byref(p))
    if rc != 0:
        raise MmsigError(rc, (lib.mmsig_last_error(None) or b"").decode())
    buf = (C.c_char * (max(n, 1) * dtype.itemsize)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)

    class _Owner:
        def __init__(self, ptr):
            self.ptr = ptr

        def __del__(self):
            try:
                lib.mmsig_host_free(C.c_void_p(self.ptr))
            except Exception:
                pass
    arr = arr.view()
    _PINNED[id(buf)] = (buf, _Owner(p.value))
    return arr


_PINNED = {}


def pack_records(term, count, out=None):
    """(term, count) int32 arrays -> the 4-byte records of mmsig_mmctm_fit_host_packed (mmsig_pack_records)."""
    lib = load()
    term = np.ascontiguousarray(term, np.int32)
    count = np.ascontiguousarray(count, np.int32)
    rec = np.empty(term.size, np.uint32) if out is None else out
    rc = lib.mmsig_pack_records(term.size, term.ctypes.data_as(c_i32p), count.ctypes.data_as(c_i32p), rec.ctypes.data_as(c_u32p))
    if rc != 0:
        raise MmsigError(rc, (lib.mmsig_last_error(None) or b"").decode())
    return rec


def comm_unique_id():
    lib = load()
    buf = (C.c_uint8 * 128)()
    rc = lib.mmsig_comm_unique_id(buf)
    if rc != 0:
        raise MmsigError(rc, (lib.mmsig_last_error(None) or b"").decode())
    return bytes(buf)
