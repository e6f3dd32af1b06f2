"""Host-side mirror of the reference's container interface, device-resident.

Same class names, method names and argument meaning as the reference's header-only containers
(``CloverVector32/4/8``, ``CloverMatrix32/4/8``: include/CloverVector4.h, include/CloverMatrix4.h, ...);
every arithmetic method forwards to the C ABI of ``libclover_b200.so`` (include/clover_b200.h), so tests
written against these classes read like the reference's own validation code
(test/validate/02_vector.cpp, 03_matrix.cpp).

PyTorch is used for plumbing only: the byte buffers are ``torch`` CUDA tensors in the reference's exact
in-memory layout (``[values | scales]`` are two tensors here), and kernels run on torch's current stream.

Differences from the reference that a caller can observe:
  * storage lives in HBM: ``getData()`` / ``getScales()`` return CUDA tensors (``.cpu().numpy()`` gives
    the reference's bytes);
  * size mismatches raise :class:`CloverSizeError` instead of printing and ``exit(1)``
    (include/CloverMatrix4.h:779-782);
  * stochastic rounding is a run-time choice: a container without a key behaves like the reference built
    with ``CLOVER_STOCHASTIC_ROUNDING_DISABLED``; ``setRandomKeys`` (include/CloverRandom.h:90-94)
    switches the reference's XORShift128+ stream on, bit-compatible with its sequential code.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import DOT_AUTO, THRESHOLD_AUTO, call

CLOVER_VECTOR_BLOCK = 64          # include/CloverVector.h:41
CLOVER_VECTOR_SIZE_PAD = 128      # include/CloverVector.h:42


class CloverSizeError(ValueError):
    """Operands do not conform (the reference prints a message and exits)."""


def size_pad(n: int) -> int:
    return n + (-n) % CLOVER_VECTOR_SIZE_PAD


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(device):
    if not torch.cuda.is_available():
        raise RuntimeError("clover_b200 needs a CUDA device: there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


class _Keyed:
    """PRNG state holder (include/CloverRandom.h): ``key`` is uint64[8] = random_key1 | random_key2."""

    key = None

    def setRandomKeys(self, key) -> None:
        if key is None:
            self.key = None
            return
        key = np.ascontiguousarray(key, dtype=np.uint64).reshape(-1)
        if key.size != 8:
            raise ValueError("a key is uint64[8] = random_key1[4] | random_key2[4]")
        self.key = key.copy()

    def seed(self, key1: int, key2: int) -> None:
        """avx_xorshift128plus_init(key1, key2, ...) (include/simdxorshift128plus.h:81-92)."""
        self.key = np.zeros(8, np.uint64)
        call("clover_prng_init", C.c_uint64(key1), C.c_uint64(key2), self.key.ctypes.data_as(C.c_void_p))

    def _key_ptr(self):
        if self.key is None:
            return None
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("a container with a PRNG key cannot be captured into a CUDA graph: the key lanes are passed by value, "
                               "every replay would reuse the same rounding noise (setRandomKeys(None) disables stochastic rounding)")
        return self.key.ctypes.data_as(C.c_void_p)

    def _generator_key(self, key):
        """The key pair a generator call consumes: an explicit uint64[8] (advanced in place, like the reference's
        `__m256i &key1, &key2` arguments) or the object's own key - seeded from the OS on first use, as the reference
        seeds every object from RDRAND (include/CloverRandom.h:96-114)."""
        if key is not None:
            if not (isinstance(key, np.ndarray) and key.dtype == np.uint64 and key.size == 8 and key.flags.c_contiguous):
                raise ValueError("a key is a contiguous uint64[8] array = random_key1[4] | random_key2[4]")
            return key
        if self.key is None:
            import os
            k = np.frombuffer(os.urandom(16), dtype=np.uint64)
            self.seed(int(k[0]) | 1, int(k[1]) | 1)
        return self.key


class _Generators(_Keyed):
    """setRandomFloats / setRandomInteger of CloverVector32 (include/CloverVector32.h:751-783, :712-744) and
    CloverMatrix32 (include/CloverMatrix32.h:252-323), run on the device, bit-identical to the reference's stream."""

    def _fill(self, fn, lo, hi, key):
        k = self._generator_key(key)
        call(fn, _ptr(self.values), C.c_uint64(self._generator_length()), C.c_float(lo), C.c_float(hi),
             k.ctypes.data_as(C.c_void_p), _stream())

    def setRandomFloats(self, min_value: float, max_value: float, key=None) -> None:
        self._fill("clover_v32_set_random_floats", min_value, max_value, key)

    def setRandomInteger(self, min_value: float, max_value: float, key=None) -> None:
        self._fill("clover_v32_set_random_integers", min_value, max_value, key)


class CloverVector32(_Generators):
    """fp32 input/output container (include/CloverVector32.h:53-70): length padded to x128, pad zeroed."""

    def _generator_length(self):
        return self.length               # the generators fill `length` elements, the pad stays 0 (:757)

    def __init__(self, n: int, data=None, device=None):
        self.length = int(n)
        self.length_pad = size_pad(self.length)
        self.values = torch.zeros(self.length_pad, dtype=torch.float32, device=_dev(device))
        if data is not None:
            src = torch.as_tensor(data, dtype=torch.float32).reshape(-1)
            if src.numel() not in (self.length, self.length_pad):
                raise CloverSizeError("data does not match the vector length")
            self.values[: src.numel()].copy_(src)

    def size(self): return self.length
    def size_pad(self): return self.length_pad
    def getBitsLength(self): return 32
    def getBytes(self): return self.length_pad * 4
    def getData(self): return self.values
    def get(self, i): return float(self.values[i])
    def numpy(self): return self.values[: self.length].cpu().numpy()


class _QVector(_Keyed):
    BITS = 0

    def __init__(self, n: int, values=None, scales=None, device=None):
        self.length = int(n)
        self.length_pad = size_pad(self.length)
        dev = _dev(device)
        nbytes = self.length_pad * self.BITS // 8
        if values is None:
            # ONE allocation [values | scales] like the reference (include/CloverVector4.h:68-103): a whole vector moves
            # between host and device with a single copy of `storage`. Pad values 0, pad scales 1 (:86-94).
            nscales = self.length_pad // 64
            self.storage = torch.zeros(nbytes + 4 * nscales, dtype=torch.uint8, device=dev)
            self.values = self.storage[:nbytes].view(torch.int8)
            self.scales = self.storage[nbytes:].view(torch.float32)
            self.scales.fill_(1.0)
        else:  # the reference's borrowing view constructor (include/CloverVector4.h:114-119)
            self.storage = None
            self.values = torch.as_tensor(values, device=dev).view(torch.int8).reshape(-1)
            self.scales = torch.as_tensor(scales, dtype=torch.float32, device=dev).reshape(-1)
            if self.values.numel() < nbytes or self.scales.numel() < self.length_pad // 64:
                raise CloverSizeError("view buffers are smaller than the padded vector")

    def size(self): return self.length
    def size_pad(self): return self.length_pad
    def getBitsLength(self): return self.BITS
    def getData(self): return self.values
    def getScales(self): return self.scales
    def getBytes(self): return self.length_pad * self.BITS // 8 + (self.length_pad // 64) * 4

    def quantize(self, other: CloverVector32) -> None:
        # the reference reads other.size_pad()/64 blocks (include/CloverVector4.h:612-613)
        if other.size_pad() != self.length_pad:
            raise CloverSizeError("Vectors do not have the same size.")
        call(f"clover_v{self.BITS}_quantize", _ptr(other.values), C.c_uint64(self.length_pad), _ptr(self.values),
             _ptr(self.scales), self._key_ptr(), _stream())

    def restore(self, other: CloverVector32) -> None:
        if other.size_pad() != self.length_pad:
            raise CloverSizeError("Vectors do not have the same size.")
        call(f"clover_v{self.BITS}_restore", _ptr(self.values), _ptr(self.scales), C.c_uint64(self.length_pad),
             _ptr(other.values), _stream())

    def dot(self, other, mode: int = DOT_AUTO) -> float:
        if other.size_pad() != self.length_pad:
            raise CloverSizeError("Vectors do not have the same size.")
        out = torch.empty(1, dtype=torch.float32, device=self.values.device)
        call(f"clover_v{self.BITS}_dot", _ptr(self.values), _ptr(self.scales), _ptr(other.values), _ptr(other.scales),
             C.c_uint64(self.length_pad), _ptr(out), C.c_int(mode), _stream())
        return float(out.item())

    def scaleAndAdd(self, other, a: float, result=None) -> None:
        """this = this + a * other (in place) or result = this + a * other, re-quantized block by block
        (include/CloverVector4.h:1195-1218, CloverVector8.h:1063-1086). Uses THIS object's PRNG key."""
        dst = self if result is None else result
        if other.size_pad() != self.length_pad or dst.size_pad() != self.length_pad:
            raise CloverSizeError("Vectors do not have the same size.")
        call(f"clover_v{self.BITS}_scale_and_add", _ptr(self.values), _ptr(self.scales), _ptr(other.values), _ptr(other.scales),
             C.c_float(a), C.c_uint64(self.length_pad), _ptr(dst.values), _ptr(dst.scales), self._key_ptr(), _stream())

    def threshold(self, k: int, mode: int = THRESHOLD_AUTO) -> None:
        """Hard thresholding in place: only the k elements of largest magnitude survive
        (include/CloverVector4.h:1913-1973, CloverVector8.h:1680-1740). mode: see clover_threshold_mode."""
        call(f"clover_v{self.BITS}_threshold", _ptr(self.values), _ptr(self.scales), C.c_uint64(self.length), C.c_uint64(int(k)),
             C.c_int(mode), _stream())

    threshold_parallel = threshold                   # include/CloverVector4.h:1919-1925: same contract, OpenMP heaps

    def clear(self) -> None:
        """all elements 0, scales 1 (include/CloverVector4.h:306-318) - the start vector of the IHT / GD loops"""
        self.values.zero_()
        self.scales.fill_(1.0)

    def dot_device(self, other, out, mode: int = DOT_AUTO) -> None:
        """dot() without the device->host read: result lands in the 1-element CUDA tensor ``out``."""
        call(f"clover_v{self.BITS}_dot", _ptr(self.values), _ptr(self.scales), _ptr(other.values), _ptr(other.scales),
             C.c_uint64(self.length_pad), _ptr(out), C.c_int(mode), _stream())

    def toVector32(self) -> CloverVector32:
        out = CloverVector32(self.length, device=self.values.device)
        self.restore(out)
        return out


class CloverVector4(_QVector):
    """4-bit vector, block-64 absmax scales (include/CloverVector4.h:44-58)."""
    BITS = 4

    def getBits(self, pos: int) -> int:               # include/CloverVector4.h:154-160
        b = int(self.values[pos >> 1])
        nib = (b >> 4) & 0xF if pos % 2 == 0 else b & 0xF
        return nib - 16 if nib >= 8 else nib

    def get(self, pos: int) -> float:                  # include/CloverVector4.h:179-188
        scale = np.float32(self.scales[pos >> 6].item()) / np.float32(7.0)
        return float(scale * np.float32(self.getBits(pos)))


class CloverVector8(_QVector):
    """8-bit vector, block-64 absmax scales (include/CloverVector8.h:45-78)."""
    BITS = 8

    def getBits(self, pos: int) -> int:
        return int(self.values[pos])

    def get(self, pos: int) -> float:                  # include/CloverVector8.h:136-139
        return float(np.float32(self.getBits(pos)) * np.float32(self.scales[pos >> 6].item()) / np.float32(127.0))


class CloverMatrix32(_Generators):
    """fp32 matrix, rows and cols padded to x128 (include/CloverMatrix.h:48-50, CloverMatrix32.h:43-50)."""

    def _generator_length(self):
        return self.rows * self.cols     # size() of the PADDED matrix: the pad is filled too (CloverMatrix32.h:254, :294)

    def __init__(self, rows: int, cols: int, data=None, device=None):
        self.rows, self.cols = size_pad(int(rows)), size_pad(int(cols))
        self.values = torch.zeros(self.rows, self.cols, dtype=torch.float32, device=_dev(device))
        if data is not None:
            src = torch.as_tensor(data, dtype=torch.float32)
            self.values[: src.shape[0], : src.shape[1]].copy_(src)

    def getRows(self): return self.rows
    def getCols(self): return self.cols
    def size(self): return self.rows * self.cols
    def getData(self): return self.values
    def getBytes(self): return self.rows * self.cols * 4


class _QMatrix(_Keyed):
    BITS = 0
    VEC = None

    def __init__(self, rows: int, cols: int, values=None, scales=None, device=None):
        self.rows, self.cols = size_pad(int(rows)), size_pad(int(cols))
        dev = _dev(device)
        nbytes = self.rows * self.cols * self.BITS // 8
        nscales = (self.rows >> 6) * (self.cols >> 6)
        if values is None:
            self.values = torch.zeros(nbytes, dtype=torch.int8, device=dev)
            self.scales = torch.zeros(nscales, dtype=torch.float32, device=dev)
        else:
            self.values = torch.as_tensor(values, device=dev).view(torch.int8).reshape(-1)
            self.scales = torch.as_tensor(scales, dtype=torch.float32, device=dev).reshape(-1)
            if self.values.numel() != nbytes or self.scales.numel() != nscales:
                raise CloverSizeError("buffers do not match the padded matrix shape")

    def getRows(self): return self.rows
    def getCols(self): return self.cols
    def size(self): return self.rows * self.cols
    def getBitsLength(self): return self.BITS
    def getData(self): return self.values
    def getScales(self): return self.scales
    def getBytes(self):                                  # include/CloverMatrix4.h:111-121
        return self.rows * self.cols * self.BITS // 8 + (self.rows >> 6) * (self.cols >> 6) * 4

    def quantize(self, m: CloverMatrix32) -> None:
        if m.getRows() != self.rows or m.getCols() != self.cols:
            raise CloverSizeError("Matrices do not have the same size.")
        call(f"clover_m{self.BITS}_quantize", _ptr(m.values), C.c_uint64(self.rows), C.c_uint64(self.cols),
             _ptr(self.values), _ptr(self.scales), self._key_ptr(), _stream())

    def mvm(self, productVector, resultVector, y32=None) -> None:
        """y = A x. (V4,V4)/(V8,V8): include/CloverMatrix4.h:777, CloverMatrix8.h:1002; (V32,V32): :1451."""
        if isinstance(productVector, CloverVector32):
            # fp32 vectors: include/CloverMatrix4.h:1451-1547, include/CloverMatrix8.h:558-661
            if productVector.size() != self.cols or resultVector.size_pad() < self.rows:
                raise CloverSizeError("MVM can not be performed.")
            call(f"clover_m{self.BITS}_mvm_f32", _ptr(self.values), _ptr(self.scales), C.c_uint64(self.rows), C.c_uint64(self.cols),
                 _ptr(productVector.values), _ptr(resultVector.values), _stream())
            return
        if self.BITS == 4 and isinstance(productVector, CloverVector8) and isinstance(resultVector, CloverVector8):
            # mixed precision (include/CloverMatrix4.h:1093-1441)
            if productVector.size() != self.cols or resultVector.size_pad() != self.rows:
                raise CloverSizeError("MVM can not be performed.")
            call("clover_m4_mvm_v8", _ptr(self.values), _ptr(self.scales), C.c_uint64(self.rows), C.c_uint64(self.cols),
                 _ptr(productVector.values), _ptr(productVector.scales), _ptr(resultVector.values), _ptr(resultVector.scales),
                 _ptr(y32), self._key_ptr(), _stream())
            return
        if not isinstance(productVector, self.VEC) or not isinstance(resultVector, self.VEC):
            raise TypeError(f"mvm expects {self.VEC.__name__} operands")
        # the reference checks productVector.size() != getCols() (include/CloverMatrix4.h:779-782)
        if productVector.size() != self.cols or resultVector.size_pad() != self.rows:
            raise CloverSizeError("MVM can not be performed.")
        call(f"clover_m{self.BITS}_mvm", _ptr(self.values), _ptr(self.scales), C.c_uint64(self.rows), C.c_uint64(self.cols),
             _ptr(productVector.values), _ptr(productVector.scales), _ptr(resultVector.values), _ptr(resultVector.scales),
             _ptr(y32), self._key_ptr(), _stream())

    def restore(self, other: CloverMatrix32) -> None:
        """other = the fp32 values this matrix represents (include/CloverMatrix4.h:266-301; 8-bit: get(i, j),
        include/CloverMatrix8.h:117-129)."""
        if other.getRows() != self.rows or other.getCols() != self.cols:
            raise CloverSizeError("Matrices do not have the same size.")
        call(f"clover_m{self.BITS}_restore", _ptr(self.values), _ptr(self.scales), C.c_uint64(self.rows), C.c_uint64(self.cols),
             _ptr(other.values), _stream())

    restore_scalar = restore

    def transpose(self, other) -> None:
        """other(j, i) = self(i, j), scales included (include/CloverMatrix4.h:1549-1663, CloverMatrix8.h:1359-1385)."""
        if type(other) is not type(self) or other.rows != self.cols or other.cols != self.rows:
            raise CloverSizeError("Matrix can not be transposed.")
        call(f"clover_m{self.BITS}_transpose", _ptr(self.values), _ptr(self.scales), C.c_uint64(self.rows),
             C.c_uint64(self.cols), _ptr(other.values), _ptr(other.scales), _stream())

    transpose_scalar = transpose_parallel = transpose


class CloverMatrix4(_QMatrix):
    """4-bit row-major matrix, one absmax scale per 64x64 tile (include/CloverMatrix4.h:38-93)."""
    BITS = 4
    VEC = CloverVector4

    def get(self, i: int, j: int) -> float:              # include/CloverMatrix4.h:123-139
        pos = i * self.cols + j
        b = int(self.values[pos >> 1])
        nib = (b >> 4) & 0xF if pos % 2 == 0 else b & 0xF
        q = nib - 16 if nib >= 8 else nib
        scale = np.float32(self.scales[(i >> 6) * (self.cols >> 6) + (j >> 6)].item()) / np.float32(7.0)
        return float(scale * np.float32(q))

    def gemm(self, Bt: "CloverMatrix4", out=None, impl: str = "tc"):
        """C = A * Bt^T with C[i][j] = rowView(A,i).dot(rowView(Bt,j)) - extension, SURVEY.md 8a-10.
        impl "tc": tcgen05 tensor-core kernel (nibbles expanded to E4M3 in a workspace first);
        impl "simt": the DP4A validation baseline. Both give the same bits."""
        if Bt.cols != self.cols:
            raise CloverSizeError("GEMM can not be performed.")
        if out is None:
            out = torch.empty(self.rows, Bt.rows, dtype=torch.float32, device=self.values.device)
        fn = {"tc": "clover_m4_gemm", "simt": "clover_m4_gemm_simt"}[impl]
        call(fn, _ptr(self.values), _ptr(self.scales), _ptr(Bt.values), _ptr(Bt.scales),
             C.c_uint64(self.rows), C.c_uint64(Bt.rows), C.c_uint64(self.cols), _ptr(out), C.c_uint64(out.stride(0)),
             _stream())
        return out

    def expand_e4m3(self, out=None):
        """rows*cols FP8-E4M3 bytes (natural element order) - the tensor-core GEMM's operand format."""
        if out is None:
            out = torch.empty(self.rows * self.cols, dtype=torch.uint8, device=self.values.device)
        call("clover_m4_expand_e4m3", _ptr(self.values), C.c_uint64(self.rows), C.c_uint64(self.cols), _ptr(out), _stream())
        return out

    def gemm_expanded(self, a8, Bt: "CloverMatrix4", bt8, out=None):
        """The GEMM on operands already expanded by expand_e4m3() (reused weights skip the expansion pass)."""
        if out is None:
            out = torch.empty(self.rows, Bt.rows, dtype=torch.float32, device=self.values.device)
        call("clover_m4_gemm_expanded", _ptr(a8), _ptr(self.scales), _ptr(bt8), _ptr(Bt.scales),
             C.c_uint64(self.rows), C.c_uint64(Bt.rows), C.c_uint64(self.cols), _ptr(out), C.c_uint64(out.stride(0)),
             _stream())
        return out


class CloverMatrix8(_QMatrix):
    """8-bit row-major matrix, one absmax scale per 64x64 tile (include/CloverMatrix8.h:76-92)."""
    BITS = 8
    VEC = CloverVector8

    def get(self, i: int, j: int) -> float:              # include/CloverMatrix8.h:117-129
        q = int(self.values[i * self.cols + j])
        scale = np.float32(self.scales[(i >> 6) * (self.cols >> 6) + (j >> 6)].item()) / np.float32(127.0)
        return float(scale * np.float32(q))
