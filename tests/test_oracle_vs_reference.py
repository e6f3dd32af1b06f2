"""Pins the C restatement (oracle/clover_oracle.c) to the UNMODIFIED reference (oracle/_ref).

Mirrors the reference's own differential validation (test/validate/02_vector.cpp,
03_matrix.cpp) but with a stricter bar: every output is compared BIT-FOR-BIT, in
rounding-disabled mode and - with an explicit PRNG key - in stochastic mode.
CPU only; runs wherever oracle/_ref was built.
"""
import numpy as np
import pytest

from oracle.pyoracle import size_pad

VEC_SIZES = [1, 63, 64, 65, 127, 128, 129, 200, 255, 256, 257, 511, 640, 1000, 1023, 1024, 4096, 8192 + 77]
MAT_SHAPES = [(128, 128), (128, 256), (256, 128), (256, 384), (384, 640), (200, 300), (640, 1152)]


def _inputs(oracle, n, kind, seed_skip=0):
    st = oracle.xs_init()
    oracle.xs_skip(st, seed_skip)
    if kind == "floats":
        return oracle.fill_floats(n, -1.0, 1.0, st)
    if kind == "ints":
        return oracle.fill_integers(n, -10.0, 10.0, st)
    if kind == "wide":
        x = oracle.fill_floats(n, -1.0, 1.0, st)
        x[:n] *= np.exp2(np.arange(n) % 40 - 20).astype(np.float32)
        return x
    raise ValueError(kind)


def test_prng_known_answer(oracle, reference):
    # lane-0 outputs of the as-written recurrence for the reference's fixed seeds
    st = oracle.xs_init()
    st_ref = reference.xs_init()
    assert np.array_equal(st, st_ref)
    expect = [0x34fdc432b801bd42, 0x17c005a764b34867, 0xd30e76c327a8ec0a, 0x678c73e394e4ac3e]
    for e in expect:
        w = oracle.xs_next(st)
        wr = reference.xs_next(st_ref)
        assert np.array_equal(w, wr)
        assert (int(w[0]) | (int(w[1]) << 32)) == e
    for _ in range(1000):
        assert np.array_equal(oracle.xs_next(st), reference.xs_next(st_ref))
    assert np.array_equal(st, st_ref)


@pytest.mark.parametrize("n", [1, 7, 8, 9, 100, 4096, 1000])
def test_generators(oracle, reference, n):
    for fn in ("fill_floats", "fill_integers"):
        st, st_ref = oracle.xs_init(), reference.xs_init()
        a = getattr(oracle, fn)(n, -1.0 if fn == "fill_floats" else -10.0, 1.0 if fn == "fill_floats" else 10.0, st)
        b = getattr(reference, fn)(n, -1.0 if fn == "fill_floats" else -10.0, 1.0 if fn == "fill_floats" else 10.0, st_ref)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        assert np.array_equal(st, st_ref)


@pytest.mark.parametrize("kind", ["floats", "ints", "wide"])
@pytest.mark.parametrize("n", VEC_SIZES)
def test_vector_quantize_restore_dot(oracle, reference, n, kind):
    x = _inputs(oracle, n, kind)
    y = _inputs(oracle, n, kind, seed_skip=3 * n + 1)
    for bits in (4, 8):
        q, r, d = (getattr(oracle, f"v{bits}_{f}") for f in ("quantize", "restore", "dot"))
        qr, rr, dr = (getattr(reference, f"v{bits}_{f}") for f in ("quantize", "restore", "dot"))
        xv, xs = q(x, n)
        xv_r, xs_r = qr(x, n)
        assert np.array_equal(xv, xv_r), f"{bits}-bit values differ"
        assert np.array_equal(xs.view(np.uint32), xs_r.view(np.uint32)), f"{bits}-bit scales differ"
        yv, ys = q(y, n)
        assert np.array_equal(r(xv, xs, n).view(np.uint32), rr(xv, xs, n).view(np.uint32))
        got, want = d(xv, xs, yv, ys, n), dr(xv, xs, yv, ys, n)
        assert got.view(np.uint32) == want.view(np.uint32), f"{bits}-bit dot {float(got).hex()} vs {float(want).hex()}"


def test_vector_zero_and_signed_zero_blocks(oracle, reference):
    n = 256
    x = np.zeros(n, np.float32)
    x[64:128] = -0.0
    x[130] = 3.5
    x[131] = -0.0
    for bits in (4, 8):
        a = getattr(oracle, f"v{bits}_quantize")(x, n)
        b = getattr(reference, f"v{bits}_quantize")(x, n)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
        assert a[1][0] == 1.0 and a[1][1] == 1.0  # zero guard: max==0 -> scale 1 (CloverVector4.h:661-663)


@pytest.mark.parametrize("n", [64, 128, 1000, 4096])
def test_vector_quantize_stochastic_with_key(oracle, reference_sr, n):
    x = _inputs(oracle, n, "floats")
    for bits in (4, 8):
        st, st_ref = oracle.xs_init(7, 9), reference_sr.xs_init(7, 9)
        a = getattr(oracle, f"v{bits}_quantize")(x, n, state=st)
        b = getattr(reference_sr, f"v{bits}_quantize")(x, n, state=st_ref)
        assert np.array_equal(a[0], b[0])
        assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
        assert np.array_equal(st, st_ref), "PRNG state after quantize differs"


@pytest.mark.parametrize("kind", ["floats", "ints"])
@pytest.mark.parametrize("shape", MAT_SHAPES)
def test_matrix_quantize_mvm(oracle, reference, shape, kind):
    from oracle.pyoracle import pad_matrix
    rows, cols = shape
    a = _inputs(oracle, rows * cols, kind)[: rows * cols].reshape(rows, cols)
    a = pad_matrix(a)
    R, Cc = a.shape
    xvec = _inputs(oracle, Cc, kind, seed_skip=11)
    for bits in (4, 8):
        mv, ms = getattr(oracle, f"m{bits}_quantize")(a)
        mv_r, ms_r, h = getattr(reference, f"m{bits}_quantize")(a)
        assert np.array_equal(mv, mv_r), f"M{bits} values"
        assert np.array_equal(ms.view(np.uint32), ms_r.view(np.uint32)), f"M{bits} scales"
        xv, xs = getattr(oracle, f"v{bits}_quantize")(xvec, Cc)
        yv, ys, y32 = getattr(oracle, f"m{bits}_mvm")(mv, ms, R, Cc, xv, xs, want_f32=True)
        yv_r, ys_r = getattr(reference, f"m{bits}_mvm")(h, xv, xs)
        assert np.array_equal(yv, yv_r), f"M{bits} mvm values"
        assert np.array_equal(ys.view(np.uint32), ys_r.view(np.uint32)), f"M{bits} mvm scales"
        # the reference's own contract: SIMD == scalar == parallel exactly (03_matrix.cpp:284, :533)
        for variant in (1, 2):
            yv_v, ys_v = getattr(reference, f"m{bits}_mvm")(h, xv, xs, variant=variant)
            assert np.array_equal(yv, yv_v) and np.array_equal(ys.view(np.uint32), ys_v.view(np.uint32))
    mv, ms = oracle.m4_quantize(a)
    _, _, h = reference.m4_quantize(a)
    y = oracle.m4_mvm_f32(mv, ms, R, Cc, xvec)
    y_r = reference.m4_mvm_f32(h, xvec)
    assert np.array_equal(y.view(np.uint32), y_r.view(np.uint32)), "mvm(V32,V32)"
    assert np.array_equal(oracle.m_restore(4, mv, ms, R, Cc).view(np.uint32), reference.m_restore(4, h).view(np.uint32)), "M4 restore"
    # round 2: CloverMatrix8::mvm(V32,V32) (CloverMatrix8.h:558-661) and the 8-bit restore (= get(i, j), :117-129)
    mv8, ms8 = oracle.m8_quantize(a)
    _, _, h8 = reference.m8_quantize(a)
    y8, y8_r = oracle.m8_mvm_f32(mv8, ms8, R, Cc, xvec), reference.m8_mvm_f32(h8, xvec)
    assert np.array_equal(y8.view(np.uint32), y8_r.view(np.uint32)), "M8 mvm(V32,V32)"
    # the reference's own acceptance bound against its double-accumulating scalar twin (03_matrix.cpp:419-491)
    assert np.max(np.abs(y8 - reference.m8_mvm_f32(h8, xvec, variant=1))) <= 0.01 * max(1.0, float(np.abs(y8).max()))
    assert np.array_equal(oracle.m_restore(8, mv8, ms8, R, Cc).view(np.uint32), reference.m_restore(8, h8).view(np.uint32)), "M8 restore"


@pytest.mark.parametrize("shape", [(128, 128), (256, 384)])
def test_matrix_stochastic_with_key(oracle, reference_sr, shape):
    from oracle.pyoracle import pad_matrix
    rows, cols = shape
    a = pad_matrix(_inputs(oracle, rows * cols, "floats")[: rows * cols].reshape(rows, cols))
    xvec = _inputs(oracle, cols, "floats", seed_skip=5)
    for bits in (4, 8):
        st, st_ref = oracle.xs_init(123, 456), reference_sr.xs_init(123, 456)
        mv, ms = getattr(oracle, f"m{bits}_quantize")(a, state=st)
        mv_r, ms_r, h = getattr(reference_sr, f"m{bits}_quantize")(a, state=st_ref)
        assert np.array_equal(mv, mv_r) and np.array_equal(ms.view(np.uint32), ms_r.view(np.uint32))
        assert np.array_equal(st, st_ref)
        xv, xs = getattr(oracle, f"v{bits}_quantize")(xvec, cols)
        yv, ys = getattr(oracle, f"m{bits}_mvm")(mv, ms, rows, cols, xv, xs, state=st)
        yv_r, ys_r = getattr(reference_sr, f"m{bits}_mvm")(h, xv, xs, state=st_ref)
        assert np.array_equal(yv, yv_r), f"M{bits} stochastic mvm values"
        assert np.array_equal(ys.view(np.uint32), ys_r.view(np.uint32))
        assert np.array_equal(st, st_ref)


def test_gemm_definition(oracle, reference):
    from oracle.pyoracle import pad_matrix
    M, N, K = 128, 256, 384
    a = pad_matrix(_inputs(oracle, M * K, "floats")[: M * K].reshape(M, K))
    b = pad_matrix(_inputs(oracle, N * K, "floats", seed_skip=999)[: N * K].reshape(N, K))
    av, as_, ha = reference.m4_quantize(a)
    bv, bs, hb = reference.m4_quantize(b)
    c = oracle.m4_gemm(av, as_, bv, bs, K, 0, M, 0, N)
    c_r = reference.m4_gemm(ha, hb, 0, M, 0, N)
    assert np.array_equal(c.view(np.uint32), c_r.view(np.uint32))
    # column j of C is mvm_f32-free: equals the fp32 intermediates of mvm(V4) with x = row j of Bt
    j = 77
    kb = K // 64
    xv = bv[j * K // 2:(j + 1) * K // 2]
    xs = bs[(j // 64) * kb:(j // 64 + 1) * kb]
    _, _, y32 = oracle.m4_mvm(av, as_, M, K, xv, xs, want_f32=True)
    assert np.array_equal(y32.view(np.uint32), c[:, j].copy().view(np.uint32))


# ---------------------------------------------------------------------------------------------------------------
# scaleAndAdd: quantized AXPY, include/CloverVector4.h:1222-1478, include/CloverVector8.h:1089-1357
# the reference's own check: test/validate/02_vector.cpp:342-394 (SIMD vs _scalar, get() by get())
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("bits", [4, 8])
@pytest.mark.parametrize("a", [0.5, -1.75, 0.0, 3.0e-3])
@pytest.mark.parametrize("n", [1, 64, 127, 128, 129, 640, 1000, 4096 + 5])
def test_scale_and_add(oracle, reference, n, a, bits):
    x, y = _inputs(oracle, n, "floats"), _inputs(oracle, n, "wide", seed_skip=321)
    q = getattr(oracle, f"v{bits}_quantize")
    (u, su), (v, sv) = q(x, n), q(y, n)
    r0, s0 = oracle.scale_and_add(bits, u, su, v, sv, a, n)
    r1, s1 = reference.scale_and_add(bits, u, su, v, sv, a, n)
    assert np.array_equal(r0, r1)
    assert np.array_equal(s0.view(np.uint32), s1.view(np.uint32))
    # in place (the reference's two-argument form passes u as the result)
    u2, su2 = u.copy(), su.copy()
    getattr(oracle.lib, f"orc_v{bits}_scale_and_add")(u2.ctypes.data_as(__import__("ctypes").c_void_p), su2.ctypes.data_as(__import__("ctypes").c_void_p),
                                                      v.ctypes.data_as(__import__("ctypes").c_void_p), sv.ctypes.data_as(__import__("ctypes").c_void_p),
                                                      __import__("ctypes").c_float(a), __import__("ctypes").c_uint64(n),
                                                      u2.ctypes.data_as(__import__("ctypes").c_void_p), su2.ctypes.data_as(__import__("ctypes").c_void_p), None)
    assert np.array_equal(u2, r1) and np.array_equal(su2.view(np.uint32), s1.view(np.uint32))


@pytest.mark.parametrize("bits", [4, 8])
@pytest.mark.parametrize("n", [128, 1000, 4096])
def test_scale_and_add_stochastic(oracle, reference_sr, n, bits):
    x, y = _inputs(oracle, n, "floats"), _inputs(oracle, n, "ints", seed_skip=77)
    q = getattr(oracle, f"v{bits}_quantize")
    (u, su), (v, sv) = q(x, n), q(y, n)
    st0, st1 = oracle.xs_init(11, 22), oracle.xs_init(11, 22)
    r0, s0 = oracle.scale_and_add(bits, u, su, v, sv, 0.5, n, st0)
    r1, s1 = reference_sr.scale_and_add(bits, u, su, v, sv, 0.5, n, st1)
    assert np.array_equal(r0, r1)
    assert np.array_equal(s0.view(np.uint32), s1.view(np.uint32))
    assert np.array_equal(st0, st1)


# ---------------------------------------------------------------------------------------------------------------
# mixed precision mvm: 4-bit matrix x CloverVector8 -> CloverVector8, include/CloverMatrix4.h:1093-1441
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["floats", "ints"])
@pytest.mark.parametrize("shape", MAT_SHAPES)
def test_matrix4_mvm_vector8(oracle, reference, shape, kind):
    from oracle.pyoracle import pad_matrix
    rows, cols = shape
    a = pad_matrix(_inputs(oracle, rows * cols, kind)[: rows * cols].reshape(rows, cols))
    R, Cc = a.shape
    xvec = _inputs(oracle, Cc, kind, seed_skip=11)
    mv, ms = oracle.m4_quantize(a)
    _, _, h = reference.m4_quantize(a)
    xv, xs = oracle.v8_quantize(xvec, Cc)
    yv, ys = oracle.m4_mvm_v8(mv, ms, R, Cc, xv, xs)
    for variant in (0, 2):                       # SIMD and _parallel
        yv_r, ys_r = reference.m4_mvm_v8(h, xv, xs, variant=variant)
        assert np.array_equal(yv, yv_r), f"values (variant {variant})"
        assert np.array_equal(ys.view(np.uint32), ys_r.view(np.uint32)), f"scales (variant {variant})"


def test_matrix4_mvm_vector8_stochastic(oracle, reference_sr):
    from oracle.pyoracle import pad_matrix
    rows, cols = 256, 384
    a = pad_matrix(_inputs(oracle, rows * cols, "floats")[: rows * cols].reshape(rows, cols))
    xvec = _inputs(oracle, cols, "floats", seed_skip=5)
    mv, ms = oracle.m4_quantize(a)
    _, _, h = reference_sr.m4_quantize(a)          # no state: rounding noise from whatever key; re-load the oracle's bytes
    import ctypes as C
    C.memmove(reference_sr.lib.ref_m4_values(h.h), mv.ctypes.data, mv.nbytes)
    C.memmove(reference_sr.lib.ref_m4_scales(h.h), ms.ctypes.data, ms.nbytes)
    xv, xs = oracle.v8_quantize(xvec, cols)
    st, st_ref = oracle.xs_init(5, 6), reference_sr.xs_init(5, 6)
    yv, ys = oracle.m4_mvm_v8(mv, ms, rows, cols, xv, xs, state=st)
    yv_r, ys_r = reference_sr.m4_mvm_v8(h, xv, xs, state=st_ref)
    assert np.array_equal(yv, yv_r) and np.array_equal(ys.view(np.uint32), ys_r.view(np.uint32))
    assert np.array_equal(st, st_ref)


@pytest.mark.parametrize("bits_", [4, 8])
@pytest.mark.parametrize("shape", [(128, 128), (128, 256), (256, 128), (384, 640), (200, 300), (640, 1152)])
def test_matrix_transpose(oracle, reference, shape, bits_):
    """transpose (03_matrix.cpp:153-246 checks get(i,j) == get(j,i)): the restatement equals every
    variant of the reference byte-for-byte - SIMD, scalar, parallel and (4-bit) scalar_faster - scales included."""
    from oracle.pyoracle import pad_matrix
    rows, cols = shape
    a = pad_matrix(_inputs(oracle, rows * cols, "floats")[: rows * cols].reshape(rows, cols))
    R, Cc = a.shape
    mv, ms, m = getattr(reference, f"m{bits_}_quantize")(a)
    tv, ts = getattr(oracle, f"m{bits_}_transpose")(mv, ms, R, Cc)
    for variant in ((0, 1, 2, 3) if bits_ == 4 else (0, 1, 2)):
        rv, rs = getattr(reference, f"m{bits_}_transpose")(m, variant)
        assert np.array_equal(tv, rv), variant
        assert np.array_equal(ts.view(np.uint32), rs.view(np.uint32)), variant
    # an involution: transposing back restores the matrix
    bv, bs = getattr(oracle, f"m{bits_}_transpose")(tv, ts, Cc, R)
    assert np.array_equal(bv, mv) and np.array_equal(bs.view(np.uint32), ms.view(np.uint32))


@pytest.mark.parametrize("bits_", [4, 8])
@pytest.mark.parametrize("n,k", [(128, 64), (129, 1), (1000, 64), (1000, 999), (1000, 1000), (2047, 64), (4096, 300), (8192 + 77, 2000)])
def test_vector_threshold(oracle, reference, n, k, bits_):
    """threshold (02_vector.cpp:450-500): the restatement - libstdc++ make_heap + the reference's
    min_heapify - leaves exactly the bytes of the reference's sequential threshold(k), ties included."""
    x = _inputs(oracle, n, "ints", seed_skip=k)          # integers in [-10, 10]: many equal magnitudes (worst case for ties)
    qv, qs = getattr(reference, f"v{bits_}_quantize")(x, n)
    want = reference.threshold(bits_, qv, qs, n, k)
    got = oracle.threshold(bits_, qv, qs, n, k)
    assert np.array_equal(got, want)
    # the reference's own acceptance test: sorted magnitudes of the survivors == the k largest magnitudes
    mags = np.sort(oracle.v_abs(bits_, qv, qs, n))[::-1]
    kept = np.sort(oracle.v_abs(bits_, got, qs, n))[::-1]
    assert np.array_equal(kept[:k], mags[:k]) and not kept[k:].any()


@pytest.mark.parametrize("bits_", [4, 8])
def test_application_loop_iht(oracle, reference, bits_):
    """The reference's IHT iteration (test/performance/01_measure.h:924-946) composed from its own sequential methods
    - mvm, scaleAndAdd (3-argument), mvm on the transpose, scaleAndAdd (in place), threshold - stays byte-identical
    between the restatement and the compiled reference over several iterations (ties in threshold included)."""
    from oracle.pyoracle import pad_matrix
    M, N, K, mu = 256, 512, 40, 0.05
    phi = pad_matrix((_inputs(oracle, M * N, "floats")[: M * N] * np.float32(0.0625)).reshape(M, N))
    y32 = _inputs(oracle, M, "floats", seed_skip=11)
    mq_o, mq_r = getattr(oracle, f"m{bits_}_quantize"), getattr(reference, f"m{bits_}_quantize")
    pv, ps = mq_o(phi)
    rpv, rps, rphi = mq_r(phi)
    assert np.array_equal(pv, rpv)
    tv, ts = getattr(oracle, f"m{bits_}_transpose")(pv, ps, M, N)
    rphit = getattr(reference, f"m{bits_}_from")(tv, ts, N, M)
    yv, ys = getattr(oracle, f"v{bits_}_quantize")(y32, M)
    mvm_o, mvm_r = getattr(oracle, f"m{bits_}_mvm"), getattr(reference, f"m{bits_}_mvm")
    xo = (np.zeros(N * bits_ // 8, np.int8), np.ones(N // 64, np.float32))
    xr = (xo[0].copy(), xo[1].copy())
    for _ in range(4):
        t1 = mvm_o(pv, ps, M, N, *xo)
        t2 = oracle.scale_and_add(bits_, yv, ys, t1[0], t1[1], -1.0, M)
        t3 = mvm_o(tv, ts, N, M, *t2)
        xo = oracle.scale_and_add(bits_, xo[0], xo[1], t3[0], t3[1], mu, N)
        xo = (oracle.threshold(bits_, xo[0], xo[1], N, K), xo[1])
        r1 = mvm_r(rphi, *xr)
        r2 = reference.scale_and_add(bits_, yv, ys, r1[0], r1[1], -1.0, M)
        r3 = mvm_r(rphit, *r2)
        xr = reference.scale_and_add(bits_, xr[0], xr[1], r3[0], r3[1], mu, N)
        xr = (reference.threshold(bits_, xr[0], xr[1], N, K), xr[1])
        assert np.array_equal(xo[0], xr[0]) and np.array_equal(xo[1].view(np.uint32), xr[1].view(np.uint32))
    assert xo[0].any()


def test_application_loop_iht_mixed(oracle, reference):
    """The reference's most accurate IHT configuration (test/accuracy/00_accuracy.cpp:84,112): 4-bit matrix, 8-bit vectors,
    mixed mvm (CloverMatrix4.h:1093-1441) - the composed iteration stays byte-identical between oracle and reference."""
    from oracle.pyoracle import pad_matrix
    M, N, K, mu = 256, 384, 30, 0.05
    phi = pad_matrix((_inputs(oracle, M * N, "floats")[: M * N] * np.float32(0.0625)).reshape(M, N))
    y32 = _inputs(oracle, M, "floats", seed_skip=11)
    pv, ps = oracle.m4_quantize(phi)
    _, _, rphi = reference.m4_quantize(phi)
    tv, ts = oracle.m4_transpose(pv, ps, M, N)
    rphit = reference.m4_from(tv, ts, N, M)
    yv, ys = oracle.v8_quantize(y32, M)
    xo = (np.zeros(N, np.int8), np.ones(N // 64, np.float32))
    xr = (xo[0].copy(), xo[1].copy())
    for _ in range(3):
        t1 = oracle.m4_mvm_v8(pv, ps, M, N, *xo)
        t2 = oracle.scale_and_add(8, yv, ys, t1[0], t1[1], -1.0, M)
        t3 = oracle.m4_mvm_v8(tv, ts, N, M, *t2)
        xo = oracle.scale_and_add(8, xo[0], xo[1], t3[0], t3[1], mu, N)
        xo = (oracle.threshold(8, xo[0], xo[1], N, K), xo[1])
        r1 = reference.m4_mvm_v8(rphi, *xr)
        r2 = reference.scale_and_add(8, yv, ys, r1[0], r1[1], -1.0, M)
        r3 = reference.m4_mvm_v8(rphit, *r2)
        xr = reference.scale_and_add(8, xr[0], xr[1], r3[0], r3[1], mu, N)
        xr = (reference.threshold(8, xr[0], xr[1], N, K), xr[1])
        assert np.array_equal(xo[0], xr[0]) and np.array_equal(xo[1].view(np.uint32), xr[1].view(np.uint32))
    assert xo[0].any()
