"""ctypes front-ends for the two CPU checkers.

TEST INFRASTRUCTURE ONLY - importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` leg. The product package ``clover_b200`` never imports it.

* :class:`Oracle`    - ``oracle/liboracle.so``, the plain-C restatement (clover_oracle.c).
* :class:`Reference` - ``oracle/_ref/libclover_ref[_sr].so``, the UNMODIFIED reference headers
  from /root/reference behind a C shim (ref_shim.cpp). Built here, shipped prebuilt to the GPU box.

Both expose the same numpy-level API so tests can swap them:
``v4_quantize, v4_restore, v4_dot, v8_*, m4_quantize, m4_mvm, m4_mvm_f32, m4_gemm, m8_quantize,
m8_mvm, xs_init, xs_next, fill_floats, fill_integers``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SEED = (445560390295639063, 2935984234003016713)  # test/random/00_random.cpp:42

_u64 = C.c_uint64
_vp = C.c_void_p


def size_pad(n: int) -> int:
    return n + (128 - n % 128) % 128


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


def aligned(n, dtype, fill=0, align=4096):
    """numpy array on a page boundary - the reference uses aligned AVX stores on its
    posix_memalign'ed buffers (e.g. _mm256_store_ps in CloverVector8.h:900-907)."""
    dtype = np.dtype(dtype)
    raw = np.empty(n * dtype.itemsize + align, np.uint8)
    off = (-raw.ctypes.data) % align
    out = raw[off:off + n * dtype.itemsize].view(dtype)
    out[...] = fill
    return out


def build(ref: bool = True) -> None:
    """(Re)build liboracle.so and, when /root/reference is present, oracle/_ref/."""
    subprocess.run(["make", "-s", "-C", HERE, "liboracle.so"] + (["ref"] if ref else []), check=True)


def padded(x: np.ndarray) -> np.ndarray:
    """fp32 vector padded with zeros to size_pad (CloverVector32 ctor, CloverVector32.h:53-70)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = aligned(size_pad(x.size), np.float32)
    out[: x.size] = x
    return out


def pad_matrix(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    r, c = a.shape
    out = aligned(size_pad(r) * size_pad(c), np.float32).reshape(size_pad(r), size_pad(c))
    out[:r, :c] = a
    return out


def v4_alloc(n):
    npad = size_pad(n)
    return aligned(npad // 2, np.int8), aligned(npad // 64, np.float32, 1)


def v8_alloc(n):
    npad = size_pad(n)
    return aligned(npad, np.int8), aligned(npad // 64, np.float32, 1)


def _state(state):
    if state is None:
        return None
    assert state.dtype == np.uint64 and state.size == 8
    return state


class Oracle:
    """The C restatement. ``state=None`` means rounding disabled; a uint64[8] state is advanced in place."""

    name = "oracle"

    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        self.lib = L = C.CDLL(path)
        L.orc_v4_dot.restype = C.c_float
        L.orc_v8_dot.restype = C.c_float
        L.orc_v4_get.restype = C.c_float
        L.orc_size_pad.restype = _u64

    # -- PRNG / generators
    def xs_init(self, k1=REF_SEED[0], k2=REF_SEED[1]):
        st = np.zeros(8, np.uint64)
        self.lib.orc_xs_init(_u64(k1), _u64(k2), _p(st))
        return st

    def xs_next(self, st):
        out = np.zeros(8, np.uint32)
        self.lib.orc_xs_next(_p(st), _p(out))
        return out

    def xs_skip(self, st, ncalls):
        self.lib.orc_xs_skip(_p(st), _u64(ncalls))

    def fill_floats(self, n, lo, hi, st):
        x = aligned(size_pad(n), np.float32)
        self.lib.orc_fill_floats(_p(x), _u64(n), C.c_float(lo), C.c_float(hi), _p(st))
        return x

    def fill_integers(self, n, lo, hi, st):
        x = aligned(size_pad(n), np.float32)
        self.lib.orc_fill_integers(_p(x), _u64(n), C.c_float(lo), C.c_float(hi), _p(st))
        return x

    # -- vectors
    def v4_quantize(self, x, n, state=None):
        v, s = v4_alloc(n)
        self.lib.orc_v4_quantize(_p(x), _u64(n), _p(v), _p(s), _p(_state(state)))
        return v, s

    def v4_restore(self, v, s, n):
        x = aligned(size_pad(n), np.float32)
        self.lib.orc_v4_restore(_p(v), _p(s), _u64(n), _p(x))
        return x

    def v4_dot(self, u, su, v, sv, n):
        return np.float32(self.lib.orc_v4_dot(_p(u), _p(su), _p(v), _p(sv), _u64(n)))

    def v8_quantize(self, x, n, state=None):
        v, s = v8_alloc(n)
        self.lib.orc_v8_quantize(_p(x), _u64(n), _p(v), _p(s), _p(_state(state)))
        return v, s

    def v8_restore(self, v, s, n):
        x = aligned(size_pad(n), np.float32)
        self.lib.orc_v8_restore(_p(v), _p(s), _u64(n), _p(x))
        return x

    def v8_dot(self, u, su, v, sv, n):
        return np.float32(self.lib.orc_v8_dot(_p(u), _p(su), _p(v), _p(sv), _u64(n)))

    def scale_and_add(self, bits, u, su, v, sv, a, n, state=None):
        """r = requantize(u + a * v) (CloverVector{4,8}::scaleAndAdd); returns (values, scales)."""
        r, sr = (v4_alloc if bits == 4 else v8_alloc)(n)
        getattr(self.lib, f"orc_v{bits}_scale_and_add")(_p(u), _p(su), _p(v), _p(sv), C.c_float(a), _u64(n), _p(r), _p(sr),
                                                        _p(_state(state)))
        return r, sr

    def threshold(self, bits, values, scales, n, k):
        """keep the k largest magnitudes (CloverVector{4,8}::threshold); returns the thresholded copy of values"""
        out = np.array(values, copy=True)
        getattr(self.lib, f"orc_v{bits}_threshold")(_p(out), _p(scales), _u64(n), _u64(k))
        return out

    def v_abs(self, bits, values, scales, n):
        """getAbs(i) for i < n (CloverVector4.h:190-203, CloverVector8.h:141-147)"""
        fn = getattr(self.lib, f"orc_v{bits}_abs")
        fn.restype = C.c_float
        return np.array([fn(_p(values), _p(scales), _u64(i)) for i in range(n)], np.float32)

    # -- matrices (a: padded fp32 [rows, cols])
    def m4_quantize(self, a, state=None):
        rows, cols = a.shape
        v = np.zeros(rows * cols // 2, np.int8)
        s = np.zeros((rows // 64) * (cols // 64), np.float32)
        self.lib.orc_m4_quantize(_p(a), _u64(rows), _u64(cols), _p(v), _p(s), _p(_state(state)))
        return v, s

    def m8_quantize(self, a, state=None):
        rows, cols = a.shape
        v = np.zeros(rows * cols, np.int8)
        s = np.zeros((rows // 64) * (cols // 64), np.float32)
        self.lib.orc_m8_quantize(_p(a), _u64(rows), _u64(cols), _p(v), _p(s), _p(_state(state)))
        return v, s

    def m4_mvm(self, mv, ms, rows, cols, xv, xs, state=None, want_f32=False):
        yv, ys = v4_alloc(rows)
        y32 = np.zeros(rows, np.float32) if want_f32 else None
        self.lib.orc_m4_mvm(_p(mv), _p(ms), _u64(rows), _u64(cols), _p(xv), _p(xs), _p(yv), _p(ys), _p(y32),
                            _p(_state(state)))
        return (yv, ys, y32) if want_f32 else (yv, ys)

    def m4_mvm_v8(self, mv, ms, rows, cols, xv, xs, state=None, want_f32=False):
        """mixed precision: 4-bit matrix x CloverVector8 -> CloverVector8 (CloverMatrix4.h:1093)"""
        yv, ys = v8_alloc(rows)
        y32 = np.zeros(rows, np.float32) if want_f32 else None
        self.lib.orc_m4_mvm_v8(_p(mv), _p(ms), _u64(rows), _u64(cols), _p(xv), _p(xs), _p(yv), _p(ys), _p(y32),
                               _p(_state(state)))
        return (yv, ys, y32) if want_f32 else (yv, ys)

    def m8_mvm(self, mv, ms, rows, cols, xv, xs, state=None, want_f32=False):
        yv, ys = v8_alloc(rows)
        y32 = np.zeros(rows, np.float32) if want_f32 else None
        self.lib.orc_m8_mvm(_p(mv), _p(ms), _u64(rows), _u64(cols), _p(xv), _p(xs), _p(yv), _p(ys), _p(y32),
                            _p(_state(state)))
        return (yv, ys, y32) if want_f32 else (yv, ys)

    def m4_mvm_f32(self, mv, ms, rows, cols, x32):
        y = np.zeros(rows, np.float32)
        self.lib.orc_m4_mvm_f32(_p(mv), _p(ms), _u64(rows), _u64(cols), _p(x32), _p(y))
        return y

    def m8_mvm_f32(self, mv, ms, rows, cols, x32):
        """CloverMatrix8::mvm(V32,V32) (CloverMatrix8.h:558-661)"""
        y = np.zeros(rows, np.float32)
        self.lib.orc_m8_mvm_f32(_p(mv), _p(ms), _u64(rows), _u64(cols), _p(x32), _p(y))
        return y

    def m_restore(self, bits, mv, ms, rows, cols):
        """matrix restore (CloverMatrix4.h:266-301; 8-bit: get(i, j), CloverMatrix8.h:117-129) -> fp32 [rows, cols]"""
        out = np.zeros((rows, cols), np.float32)
        getattr(self.lib, f"orc_m{bits}_restore")(_p(mv), _p(ms), _u64(rows), _u64(cols), _p(out))
        return out

    def m4_gemm(self, av, as_, btv, bts, K, i0, i1, j0, j1):
        c = np.zeros((i1 - i0, j1 - j0), np.float32)
        self.lib.orc_m4_gemm(_p(av), _p(as_), _p(btv), _p(bts), _u64(K), _u64(i0), _u64(i1), _u64(j0), _u64(j1),
                             _p(c), _u64(j1 - j0))
        return c

    def m4_transpose(self, mv, ms, rows, cols):
        """(values, scales) of the cols x rows transpose (CloverMatrix4.h:1549-1663)"""
        ov = np.zeros(rows * cols // 2, np.int8)
        os_ = np.zeros((rows // 64) * (cols // 64), np.float32)
        self.lib.orc_m4_transpose(_p(mv), _p(ms), _u64(rows), _u64(cols), _p(ov), _p(os_))
        return ov, os_

    def m8_transpose(self, mv, ms, rows, cols):
        ov = np.zeros(rows * cols, np.int8)
        os_ = np.zeros((rows // 64) * (cols // 64), np.float32)
        self.lib.orc_m8_transpose(_p(mv), _p(ms), _u64(rows), _u64(cols), _p(ov), _p(os_))
        return ov, os_


class Reference:
    """The unmodified reference (oracle/_ref). ``stochastic`` picks the build flavour.

    ``variant``: 0 = SIMD (unsuffixed methods), 1 = ``_scalar``, 2 = ``_parallel``.
    """

    name = "reference"

    def __init__(self, stochastic: bool = False, threads: int = 0):
        """threads > 0: size the reference's OpenMP team explicitly (only effective in the first Reference of a process,
        the reference reads its team size once - include/CloverBase.h:369-380)."""
        fn = "libclover_ref_sr.so" if stochastic else "libclover_ref.so"
        path = os.path.join(HERE, "_ref", fn)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        self.lib = L = C.CDLL(path)
        self.stochastic = stochastic
        for f in ("ref_v4_dot", "ref_v8_dot", "ref_v4_get", "ref_m4_get"):
            getattr(L, f).restype = C.c_float
        for f in ("ref_m32_create", "ref_m4_create", "ref_m8_create", "ref_m32_data", "ref_m4_values",
                  "ref_m4_scales", "ref_m8_values", "ref_m8_scales"):
            getattr(L, f).restype = _vp
        for f in ("ref_m32_rows", "ref_m32_cols", "ref_m4_rows", "ref_m4_cols", "ref_m4_bytes", "ref_m8_rows",
                  "ref_m8_cols", "ref_m8_bytes", "ref_v4_bytes", "ref_v_size_pad"):
            getattr(L, f).restype = _u64
        assert bool(L.ref_stochastic_enabled()) == stochastic
        if threads > 0:
            L.ref_set_openmp_threads(C.c_int(threads))

    @staticmethod
    def available(stochastic: bool = False) -> bool:
        fn = "libclover_ref_sr.so" if stochastic else "libclover_ref.so"
        return os.path.exists(os.path.join(HERE, "_ref", fn))

    def threads(self):
        return int(self.lib.ref_openmp_threads())

    def _st(self, state):
        # the rounding-disabled build ignores keys; the stochastic build needs one to be deterministic
        return _p(_state(state))

    def xs_init(self, k1=REF_SEED[0], k2=REF_SEED[1]):
        st = np.zeros(8, np.uint64)
        self.lib.ref_xs_init(_u64(k1), _u64(k2), _p(st))
        return st

    def xs_next(self, st):
        out = np.zeros(8, np.uint32)
        self.lib.ref_xs_next(_p(st), _p(out))
        return out

    def fill_floats(self, n, lo, hi, st):
        x = aligned(size_pad(n), np.float32)
        self.lib.ref_fill_floats(_p(x), _u64(n), C.c_float(lo), C.c_float(hi), _p(st))
        return x

    def fill_integers(self, n, lo, hi, st):
        x = aligned(size_pad(n), np.float32)
        self.lib.ref_fill_integers(_p(x), _u64(n), C.c_float(lo), C.c_float(hi), _p(st))
        return x

    def v4_quantize(self, x, n, state=None, variant=0):
        v, s = v4_alloc(n)
        self.lib.ref_v4_quantize(_p(x), _u64(n), _p(v), _p(s), self._st(state), C.c_int(variant))
        return v, s

    def v4_restore(self, v, s, n, variant=0):
        x = aligned(size_pad(n), np.float32)
        self.lib.ref_v4_restore(_p(v), _p(s), _u64(n), _p(x), C.c_int(variant))
        return x

    def v4_dot(self, u, su, v, sv, n, variant=0):
        return np.float32(self.lib.ref_v4_dot(_p(u), _p(su), _p(v), _p(sv), _u64(n), C.c_int(variant)))

    def v4_get(self, v, s, n, i):
        return np.float32(self.lib.ref_v4_get(_p(v), _p(s), _u64(n), _u64(i)))

    def v8_quantize(self, x, n, state=None, variant=0):
        v, s = v8_alloc(n)
        self.lib.ref_v8_quantize(_p(x), _u64(n), _p(v), _p(s), self._st(state), C.c_int(variant))
        return v, s

    def v8_restore(self, v, s, n, variant=0):
        x = aligned(size_pad(n), np.float32)
        self.lib.ref_v8_restore(_p(v), _p(s), _u64(n), _p(x), C.c_int(variant))
        return x

    def v8_dot(self, u, su, v, sv, n, variant=0):
        return np.float32(self.lib.ref_v8_dot(_p(u), _p(su), _p(v), _p(sv), _u64(n), C.c_int(variant)))

    def scale_and_add(self, bits, u, su, v, sv, a, n, state=None, variant=0):
        r, sr = (v4_alloc if bits == 4 else v8_alloc)(n)
        getattr(self.lib, f"ref_v{bits}_scale_and_add")(_p(u), _p(su), _p(v), _p(sv), C.c_float(a), _u64(n), _p(r), _p(sr),
                                                        self._st(state), C.c_int(variant))
        return r, sr

    def threshold(self, bits, values, scales, n, k, variant=0):
        out = np.array(values, copy=True)
        sc = np.array(scales, copy=True)
        getattr(self.lib, f"ref_v{bits}_threshold")(_p(out), _p(sc), _u64(n), _u64(k), C.c_int(variant))
        return out

    # -- matrices: handle based (the reference's matrices own their storage)
    class _M:
        def __init__(self, ref, bits, rows, cols):
            self.ref, self.bits = ref, bits
            self.L = ref.lib
            self.pfx = f"ref_m{bits}_"
            self.h = _vp(getattr(self.L, self.pfx + "create")(_u64(rows), _u64(cols)))
            self.rows = int(getattr(self.L, self.pfx + "rows")(self.h))
            self.cols = int(getattr(self.L, self.pfx + "cols")(self.h))

        def __del__(self):
            if getattr(self, "h", None):
                getattr(self.L, self.pfx + "destroy")(self.h)
                self.h = None

        def _view(self, ptr, nbytes, dtype):
            buf = (C.c_char * nbytes).from_address(ptr)
            return np.frombuffer(buf, dtype=dtype)

        @property
        def values(self):
            nbytes = self.rows * self.cols * (4 if self.bits == 32 else 1) // (2 if self.bits == 4 else 1)
            fn = "data" if self.bits == 32 else "values"
            return self._view(getattr(self.L, self.pfx + fn)(self.h), nbytes, np.float32 if self.bits == 32 else np.int8)

        @property
        def scales(self):
            n = (self.rows // 64) * (self.cols // 64)
            return self._view(getattr(self.L, self.pfx + "scales")(self.h), 4 * n, np.float32)

    def m32(self, a_padded):
        rows, cols = a_padded.shape
        m = Reference._M(self, 32, rows, cols)
        m.values[:] = a_padded.reshape(-1)
        return m

    def m4_quantize(self, a, state=None, variant=0):
        """a: padded fp32 [rows, cols] -> (values, scales, handle)"""
        src = self.m32(a)
        m = Reference._M(self, 4, *a.shape)
        self.lib.ref_m4_quantize(m.h, src.h, self._st(state), C.c_int(variant))
        return m.values.copy(), m.scales.copy(), m

    def m8_quantize(self, a, state=None, variant=0):
        src = self.m32(a)
        m = Reference._M(self, 8, *a.shape)
        self.lib.ref_m8_quantize(m.h, src.h, self._st(state), C.c_int(variant))
        return m.values.copy(), m.scales.copy(), m

    def m4_from(self, mv, ms, rows, cols):
        m = Reference._M(self, 4, rows, cols)
        m.values[:] = mv
        m.scales[:] = ms
        return m

    def m8_from(self, mv, ms, rows, cols):
        m = Reference._M(self, 8, rows, cols)
        m.values[:] = mv
        m.scales[:] = ms
        return m

    def m4_mvm(self, m, xv, xs, state=None, variant=0):
        yv, ys = v4_alloc(m.rows)
        self.lib.ref_m4_mvm(m.h, _p(xv), _p(xs), _p(yv), _p(ys), self._st(state), C.c_int(variant))
        return yv, ys

    def m4_mvm_v8(self, m, xv, xs, state=None, variant=0):
        yv, ys = v8_alloc(m.rows)
        self.lib.ref_m4_mvm_v8(m.h, _p(xv), _p(xs), _p(yv), _p(ys), self._st(state), C.c_int(variant))
        return yv, ys

    def m8_mvm(self, m, xv, xs, state=None, variant=0):
        yv, ys = v8_alloc(m.rows)
        self.lib.ref_m8_mvm(m.h, _p(xv), _p(xs), _p(yv), _p(ys), self._st(state), C.c_int(variant))
        return yv, ys

    def m4_transpose(self, m, variant=0):
        """variant: 0 SIMD, 1 scalar, 2 parallel, 3 scalar_faster"""
        o = Reference._M(self, 4, m.cols, m.rows)
        self.lib.ref_m4_transpose(m.h, o.h, C.c_int(variant))
        return o.values.copy(), o.scales.copy()

    def m8_transpose(self, m, variant=0):
        o = Reference._M(self, 8, m.cols, m.rows)
        self.lib.ref_m8_transpose(m.h, o.h, C.c_int(variant))
        return o.values.copy(), o.scales.copy()

    def m4_mvm_f32(self, m, x32, variant=0):
        y = aligned(size_pad(m.rows), np.float32)
        self.lib.ref_m4_mvm_f32(m.h, _p(x32), _p(y), C.c_int(variant))
        return y[: m.rows]

    def m8_mvm_f32(self, m, x32, variant=0):
        y = aligned(size_pad(m.rows), np.float32)
        self.lib.ref_m8_mvm_f32(m.h, _p(x32), _p(y), C.c_int(variant))
        return y[: m.rows]

    def m_restore(self, bits, m):
        out = np.zeros((m.rows, m.cols), np.float32)
        if bits == 4:
            self.lib.ref_m4_restore(m.h, _p(out))
        else:
            self.lib.ref_m8_restore_by_get(m.h, _p(out))
        return out

    def m4_gemm(self, ma, mbt, i0, i1, j0, j1):
        c = np.zeros((i1 - i0, j1 - j0), np.float32)
        self.lib.ref_m4_gemm_rows(ma.h, mbt.h, _u64(i0), _u64(i1), _u64(j0), _u64(j1), _p(c), _u64(j1 - j0))
        return c


def fnv1a64(data) -> str:
    """64-bit FNV-1a over raw bytes (the hash SURVEY.md 8c's known-answer table uses)."""
    h = 1469598103934665603
    for b in bytes(data):
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"
