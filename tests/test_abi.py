"""The C-ABI shared library loads and exports exactly what include/clover_b200.h declares (no GPU needed),
and the host-only entry points (padding rule, PRNG init/next/jump-ahead) agree with the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "clover_b200.h")


@pytest.fixture(scope="module")
def lib():
    import clover_b200
    clover_b200.build()
    return clover_b200.lib()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(clover_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_python_binds():
    import clover_b200
    assert declared_symbols() == sorted(clover_b200.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", lib._name], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (clover_[a-z0-9_]+)", out))
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, f"declared in include/clover_b200.h but not exported: {missing}"
    assert all(s.startswith("clover_") for s in exported)


def test_library_is_sm100a_and_carries_no_oracle(lib):
    sass = subprocess.run(["cuobjdump", "-lelf", lib._name], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    needed = subprocess.run(["readelf", "-d", lib._name], capture_output=True, text=True).stdout
    assert "oracle" not in needed and "clover_ref" not in needed


def test_size_pad_rule(lib):
    for n in (0, 1, 127, 128, 129, 1000, 4096, (1 << 26) + 1):
        assert lib.clover_size_pad(n) == n + (-n) % 128


def test_compute_fails_loudly_without_gpu(lib):
    import ctypes as C
    import clover_b200
    if lib.clover_device_count() > 0:
        pytest.skip("a GPU is present")
    buf = (C.c_char * 4096)()
    with pytest.raises(clover_b200.CloverError):
        clover_b200.call("clover_v4_quantize", C.cast(buf, C.c_void_p), C.c_uint64(128), C.cast(buf, C.c_void_p),
                         C.cast(buf, C.c_void_p), None, None)


def test_host_prng_matches_oracle(lib, oracle):
    import ctypes as C
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    key = np.zeros(8, np.uint64)
    assert lib.clover_prng_init(445560390295639063, 2935984234003016713, vp(key)) == 0
    st = oracle.xs_init()
    assert np.array_equal(key, st)
    out = np.zeros(8, np.uint32)
    for _ in range(100):
        lib.clover_prng_next(vp(key), vp(out))
        assert np.array_equal(out, oracle.xs_next(st))
    assert np.array_equal(key, st)
    # O(log n) jump-ahead == n sequential calls, including the trailing part1 lanes
    for n in (1, 2, 3, 64, 1000, 12345, 2 * (1 << 20)):
        a, b = key.copy(), st.copy()
        assert lib.clover_prng_skip(vp(a), n) == 0
        oracle.xs_skip(b, n)
        assert np.array_equal(a, b), n
