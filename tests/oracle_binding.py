"""ctypes bindings for the CHECKERS under oracle/ (test infrastructure).

  Oracle  -- oracle/liboracle.so, the plain-C restatement (always available
             after `make -C oracle`; built on demand here).
  Ref     -- oracle/_ref/libspmv_ref.so, the unmodified reference sources
             compiled with oracle/ref_shim.cpp (present when it was built in
             the authoring container; optional).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import
this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libspmv_ref.so")

fp = C.POINTER(C.c_float)
ip = C.POINTER(C.c_int)
dp = C.POINTER(C.c_double)


def _f(a):
    return a.ctypes.data_as(fp)


def _i(a):
    return a.ctypes.data_as(ip)


def _d(a):
    return a.ctypes.data_as(dp)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Stats(C.Structure):
    _fields_ = [("avg", C.c_float), ("max", C.c_int), ("min", C.c_int), ("skew", C.c_float)]


class TopK(C.Structure):
    _fields_ = [("node_id", C.c_int), ("rank", C.c_float)]


class Oracle:
    """The C restatement.  Methods take/return numpy arrays."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
                os.path.join(ORACLE_DIR, "spmv_oracle.c")):
            subprocess.check_call(["make", "-C", ORACLE_DIR, os.path.join(ORACLE_DIR, "liboracle.so")],
                                  stdout=subprocess.DEVNULL)
        L = self.L = C.CDLL(ORACLE_SO)
        L.orc_csr_get_element.restype = C.c_float
        L.orc_ell_get_element.restype = C.c_float
        L.orc_bytes_csr.restype = C.c_uint64
        L.orc_bytes_ell.restype = C.c_uint64
        L.orc_achieved_gbs.restype = C.c_float
        L.orc_achieved_gbs.argtypes = [C.c_uint64, C.c_float]
        L.orc_fnv1a64.restype = C.c_uint64
        L.orc_fnv1a64.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_pagerank_f32.argtypes = [C.c_int, C.c_int, fp, ip, ip, C.c_float, C.c_float, C.c_int, fp, fp, ip]
        L.orc_pagerank_f64.argtypes = [C.c_int, C.c_int, fp, ip, ip, C.c_float, C.c_float, C.c_int, C.c_int, fp, dp,
                                       dp, ip]

    # -- formats
    def csr_from_dense(self, dense):
        dense = f32(dense)
        rows, cols = dense.shape
        nnz = self.L.orc_dense_count_nnz(_f(dense), rows, cols)
        va, ci, rp = np.zeros(nnz, np.float32), np.zeros(nnz, np.int32), np.zeros(rows + 1, np.int32)
        self.L.orc_csr_from_dense(_f(dense), rows, cols, _f(va), _i(ci), _i(rp))
        return rp, ci, va

    def csr_to_dense(self, rows, cols, rp, ci, va):
        out = np.empty((rows, cols), np.float32)
        self.L.orc_csr_to_dense(rows, cols, _f(f32(va)), _i(i32(ci)), _i(i32(rp)), _f(out))
        return out

    def csr_get_element(self, rows, cols, rp, ci, va, r, c):
        return self.L.orc_csr_get_element(rows, cols, _f(f32(va)), _i(i32(ci)), _i(i32(rp)), r, c)

    def stats(self, rows, nnz, rp):
        s = Stats()
        self.L.orc_csr_stats(rows, nnz, _i(i32(rp)), C.byref(s))
        return s

    def auto_config(self, rows, cols, nnz, rp):
        bs, tex = C.c_int(0), C.c_int(0)
        kt = self.L.orc_auto_config(rows, cols, nnz, _i(i32(rp)), C.byref(bs), C.byref(tex))
        return kt, bs.value, bool(tex.value)

    def ell_from_csr(self, rows, rp, ci, va):
        rp, ci, va = i32(rp), i32(ci), f32(va)
        w = self.L.orc_ell_width_from_csr(rows, _i(rp))
        ev, ec = np.zeros(rows * w, np.float32), np.zeros(rows * w, np.int32)
        self.L.orc_ell_from_csr(rows, w, _f(va), _i(ci), _i(rp), _f(ev), _i(ec))
        return w, ec, ev

    def ell_from_dense(self, dense):
        dense = f32(dense)
        rows, cols = dense.shape
        w = self.L.orc_ell_width_from_dense(_f(dense), rows, cols)
        ev, ec = np.zeros(rows * w, np.float32), np.zeros(rows * w, np.int32)
        self.L.orc_ell_from_dense(_f(dense), rows, cols, w, _f(ev), _i(ec))
        return w, ec, ev

    def ell_to_dense(self, rows, cols, w, ec, ev):
        out = np.empty((rows, cols), np.float32)
        self.L.orc_ell_to_dense(rows, cols, w, _f(f32(ev)), _i(i32(ec)), _f(out))
        return out

    def ell_get_element(self, rows, cols, w, ec, ev, r, c):
        return self.L.orc_ell_get_element(rows, cols, w, _f(f32(ev)), _i(i32(ec)), r, c)

    # -- SpMV
    def spmv_csr(self, rows, rp, ci, va, x):
        y = np.empty(rows, np.float32)
        self.L.orc_spmv_csr(rows, _f(f32(va)), _i(i32(ci)), _i(i32(rp)), _f(f32(x)), _f(y))
        return y

    def spmv_csr_f64(self, rows, rp, ci, va, x):
        y, s = np.empty(rows, np.float64), np.empty(rows, np.float64)
        self.L.orc_spmv_csr_f64(rows, _f(f32(va)), _i(i32(ci)), _i(i32(rp)), _f(f32(x)), _d(y), _d(s))
        return y, s

    def spmv_ell(self, rows, w, ec, ev, x):
        y = np.empty(rows, np.float32)
        self.L.orc_spmv_ell(rows, w, _f(f32(ev)), _i(i32(ec)), _f(f32(x)), _f(y))
        return y

    def spmv_ell_f64(self, rows, w, ec, ev, x):
        y, s = np.empty(rows, np.float64), np.empty(rows, np.float64)
        self.L.orc_spmv_ell_f64(rows, w, _f(f32(ev)), _i(i32(ec)), _f(f32(x)), _d(y), _d(s))
        return y, s

    # -- bandwidth
    def bytes_csr(self, rows, cols, nnz):
        return int(self.L.orc_bytes_csr(rows, cols, nnz))

    def bytes_ell(self, rows, cols, w):
        return int(self.L.orc_bytes_ell(rows, cols, w))

    def achieved_gbs(self, nbytes, ms):
        return self.L.orc_achieved_gbs(nbytes, ms)

    # -- PageRank
    def find_dangling(self, rows, cols, rp, ci, va):
        flags = np.zeros(max(cols, 1), np.uint8)
        n = self.L.orc_find_dangling(rows, cols, _f(f32(va)), _i(i32(ci)), _i(i32(rp)), flags.ctypes.data_as(C.c_void_p))
        return n, flags[:cols]

    def pagerank_f32(self, n, cols, rp, ci, va, damping=0.85, tol=1e-6, max_it=100):
        ranks = np.empty(n, np.float32)
        res, conv = C.c_float(0), C.c_int(0)
        it = self.L.orc_pagerank_f32(n, cols, _f(f32(va)), _i(i32(ci)), _i(i32(rp)), damping, tol, max_it, _f(ranks),
                                     C.byref(res), C.byref(conv))
        return ranks, it, res.value, bool(conv.value)

    def pagerank_f64(self, n, cols, rp, ci, va, damping=0.85, tol=1e-6, max_it=100, fixed_it=0):
        ranks = np.empty(n, np.float32)
        l2, l1, conv = C.c_double(0), C.c_double(0), C.c_int(0)
        it = self.L.orc_pagerank_f64(n, cols, _f(f32(va)), _i(i32(ci)), _i(i32(rp)), damping, tol, max_it, fixed_it,
                                     _f(ranks), C.byref(l2), C.byref(l1), C.byref(conv))
        return ranks, it, l2.value, l1.value, bool(conv.value)

    def top_k(self, ranks, k):
        ranks = f32(ranks)
        kk = min(k, len(ranks))
        out = (TopK * max(kk, 1))()
        self.L.orc_top_k(_f(ranks), len(ranks), k, out)
        return (np.array([out[i].node_id for i in range(kk)], np.int32),
                np.array([out[i].rank for i in range(kk)], np.float32))

    # -- merge path / partition
    def merge_path_search(self, diagonal, rp, rows, nnz):
        r, z = C.c_int(0), C.c_int(0)
        self.L.orc_merge_path_search(diagonal, _i(i32(rp)), rows, nnz, C.byref(r), C.byref(z))
        return r.value, z.value

    def partition_rows(self, rows, rp, parts, row_weight=0):
        b = np.zeros(parts + 1, np.int32)
        self.L.orc_partition_rows_weighted(rows, _i(i32(rp)), parts, row_weight, _i(b))
        return b

    def fnv(self, arr):
        arr = np.ascontiguousarray(arr)
        return int(self.L.orc_fnv1a64(arr.ctypes.data_as(C.c_void_p), arr.nbytes))


class Ref:
    """The unmodified reference library behind oracle/ref_shim.cpp."""

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def __init__(self):
        L = self.L = C.CDLL(REF_SO)
        for name in ("ref_csr_wrap", "ref_csr_from_dense", "ref_ell_from_dense", "ref_ell_from_csr"):
            getattr(L, name).restype = C.c_void_p
        L.ref_csr_wrap.argtypes = [C.c_int, C.c_int, C.c_int, fp, ip, ip]
        L.ref_csr_from_dense.argtypes = [fp, C.c_int, C.c_int, ip]
        L.ref_ell_from_dense.argtypes = [fp, C.c_int, C.c_int, ip]
        L.ref_ell_from_csr.argtypes = [C.c_void_p, ip]
        for name in ("ref_csr_destroy", "ref_ell_destroy"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = None
        L.ref_csr_fields.argtypes = [C.c_void_p, ip, ip, ip, C.POINTER(fp), C.POINTER(ip), C.POINTER(ip)]
        L.ref_ell_fields.argtypes = [C.c_void_p, ip, ip, ip, C.POINTER(fp), C.POINTER(ip)]
        L.ref_csr_get_element.restype = C.c_float
        L.ref_csr_get_element.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_ell_get_element.restype = C.c_float
        L.ref_ell_get_element.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_csr_to_dense.argtypes = [C.c_void_p, fp]
        L.ref_ell_to_dense.argtypes = [C.c_void_p, fp]
        L.ref_csr_to_gpu.argtypes = [C.c_void_p]
        L.ref_ell_to_gpu.argtypes = [C.c_void_p]
        L.ref_csr_serialize.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_csr_deserialize.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_ell_serialize.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_ell_deserialize.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_csr_stats.argtypes = [C.c_void_p, fp, ip, ip, fp]
        L.ref_auto_config.argtypes = [C.c_void_p, ip, ip, ip]
        L.ref_spmv_cpu_csr.argtypes = [C.c_void_p, fp, fp]
        L.ref_spmv_cpu_ell.argtypes = [C.c_void_p, fp, fp]
        L.ref_time_spmv_cpu_csr.restype = C.c_double
        L.ref_time_spmv_cpu_csr.argtypes = [C.c_void_p, fp, fp, C.c_int]
        L.ref_time_spmv_cpu_ell.restype = C.c_double
        L.ref_time_spmv_cpu_ell.argtypes = [C.c_void_p, fp, fp, C.c_int]
        L.ref_spmv_csr_gpu.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, fp, fp, fp]
        L.ref_spmv_ell_gpu.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, fp, fp, fp]
        L.ref_pagerank.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_int, fp, fp, ip]
        L.ref_pagerank_top_k.argtypes = [fp, C.c_int, C.c_int, ip, fp]
        L.ref_bandwidth_csr.argtypes = [C.c_void_p, C.c_float, fp, fp, fp]
        L.ref_bandwidth_ell.argtypes = [C.c_void_p, C.c_float, fp, fp, fp]
        L.ref_benchmark_to_json.argtypes = [C.c_char_p, fp, C.c_int, C.c_char_p, C.c_int]
        L.ref_benchmark_from_json.argtypes = [C.c_char_p, fp, ip]
        L.ref_rng_seed.argtypes = [C.c_uint]
        L.ref_rng_int.argtypes = [C.c_int, C.c_int]
        L.ref_rng_float.restype = C.c_float
        L.ref_rng_float.argtypes = [C.c_float, C.c_float]
        L.ref_rng_dense.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, fp]
        L.ref_rng_vector.argtypes = [C.c_int, C.c_float, C.c_float, fp]

    # handles -----------------------------------------------------------------
    def csr_wrap(self, rows, cols, rp, ci, va):
        rp, ci, va = i32(rp), i32(ci), f32(va)
        h = self.L.ref_csr_wrap(rows, cols, len(va), _f(va), _i(ci), _i(rp))
        return h, (rp, ci, va)  # keep arrays alive with the handle

    def csr_from_dense(self, dense):
        dense = f32(dense)
        st = C.c_int(0)
        h = self.L.ref_csr_from_dense(_f(dense), dense.shape[0], dense.shape[1], C.byref(st))
        return h, st.value

    def csr_fields(self, h):
        r, c, n = C.c_int(), C.c_int(), C.c_int()
        va, ci, rp = fp(), ip(), ip()
        self.L.ref_csr_fields(h, C.byref(r), C.byref(c), C.byref(n), C.byref(va), C.byref(ci), C.byref(rp))
        rows, cols, nnz = r.value, c.value, n.value
        rp_a = np.ctypeslib.as_array(rp, shape=(rows + 1,)).copy()
        ci_a = np.ctypeslib.as_array(ci, shape=(nnz,)).copy() if nnz else np.zeros(0, np.int32)
        va_a = np.ctypeslib.as_array(va, shape=(nnz,)).copy() if nnz else np.zeros(0, np.float32)
        return rows, cols, nnz, rp_a, ci_a, va_a

    def ell_fields(self, h):
        r, c, w = C.c_int(), C.c_int(), C.c_int()
        va, ci = fp(), ip()
        self.L.ref_ell_fields(h, C.byref(r), C.byref(c), C.byref(w), C.byref(va), C.byref(ci))
        n = r.value * w.value
        ci_a = np.ctypeslib.as_array(ci, shape=(n,)).copy() if n else np.zeros(0, np.int32)
        va_a = np.ctypeslib.as_array(va, shape=(n,)).copy() if n else np.zeros(0, np.float32)
        return r.value, c.value, w.value, ci_a, va_a

    def ell_from_dense(self, dense):
        dense = f32(dense)
        st = C.c_int(0)
        return self.L.ref_ell_from_dense(_f(dense), dense.shape[0], dense.shape[1], C.byref(st)), st.value

    def ell_from_csr(self, h):
        st = C.c_int(0)
        return self.L.ref_ell_from_csr(h, C.byref(st)), st.value

    def stats(self, h):
        a, s = C.c_float(), C.c_float()
        mx, mn = C.c_int(), C.c_int()
        self.L.ref_csr_stats(h, C.byref(a), C.byref(mx), C.byref(mn), C.byref(s))
        return a.value, mx.value, mn.value, s.value

    def auto_config(self, h):
        k, b, t = C.c_int(), C.c_int(), C.c_int()
        self.L.ref_auto_config(h, C.byref(k), C.byref(b), C.byref(t))
        return k.value, b.value, bool(t.value)

    def spmv_cpu_csr(self, h, x, rows):
        y = np.empty(rows, np.float32)
        self.L.ref_spmv_cpu_csr(h, _f(f32(x)), _f(y))
        return y

    def spmv_cpu_ell(self, h, x, rows):
        y = np.empty(rows, np.float32)
        self.L.ref_spmv_cpu_ell(h, _f(f32(x)), _f(y))
        return y

    # replay of the reference test fixtures -------------------------------------
    def rng_seed(self, seed):
        self.L.ref_rng_seed(seed)

    def rng_int(self, lo, hi):
        return self.L.ref_rng_int(lo, hi)

    def rng_float(self, lo, hi):
        return self.L.ref_rng_float(lo, hi)

    def rng_dense(self, rows, cols, density, lo=-10.0, hi=10.0):
        out = np.empty((rows, cols), np.float32)
        self.L.ref_rng_dense(rows, cols, density, lo, hi, _f(out))
        return out

    def rng_vector(self, n, lo=-10.0, hi=10.0):
        out = np.empty(n, np.float32)
        self.L.ref_rng_vector(n, lo, hi, _f(out))
        return out
