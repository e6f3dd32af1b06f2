"""Parity of the CUDA path (through the C ABI) against the CPU oracle. Needs a B200.

Structure follows the reference's own validation (test/validate/02_vector.cpp: quantization :111-144,
restore :223-256, dot :258-295; test/validate/03_matrix.cpp: quantization :38-96, MVM :248-326,
mvm(V32) :419-491) with a stricter bar: the reference compares `get(i)` of two implementations, these
tests compare every packed byte and every fp32 scale BIT-FOR-BIT.

The oracle (oracle/clover_oracle.c) is itself pinned bit-for-bit to the unmodified reference by
tests/test_oracle_vs_reference.py and tests/test_golden.py.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def cb():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import clover_b200
    clover_b200.lib()
    from clover_b200 import containers
    return containers


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def gen(oracle, n, kind, skip=0):
    st = oracle.xs_init()
    oracle.xs_skip(st, skip)
    if kind == "ints":
        return oracle.fill_integers(n, -10.0, 10.0, st)
    x = oracle.fill_floats(n, -1.0, 1.0, st)
    if kind == "wide":
        x[:n] *= np.exp2(np.arange(n) % 40 - 20).astype(np.float32)
    return x


# the reference sweeps 128..1023 step 1 (02_vector.cpp:118); every residue class mod 128 behaves the
# same after padding, so sample it and add the large / ragged cases
VEC_SIZES = [1, 64, 127, 128, 129, 255, 640, 1000, 1023, 4096, 65536 + 3, (1 << 20) + 128]


@pytest.mark.parametrize("bits_", [4, 8])
@pytest.mark.parametrize("kind", ["floats", "ints", "wide"])
@pytest.mark.parametrize("n", VEC_SIZES)
def test_vector_quantize_restore(cb, oracle, n, kind, bits_):
    V = cb.CloverVector4 if bits_ == 4 else cb.CloverVector8
    x = gen(oracle, n, kind)
    q = V(n)
    q.quantize(cb.CloverVector32(n, x))
    ov, os_ = getattr(oracle, f"v{bits_}_quantize")(x, n)
    assert np.array_equal(q.getData().cpu().numpy(), ov)
    assert np.array_equal(bits(q.getScales().cpu().numpy()), bits(os_))
    r = cb.CloverVector32(n)
    q.restore(r)
    assert np.array_equal(bits(r.getData().cpu().numpy()), bits(getattr(oracle, f"v{bits_}_restore")(ov, os_, n)))
    assert q.getBytes() == ov.nbytes + os_.nbytes
    for i in (0, n // 2, n - 1):
        if bits_ == 4:      # get(): (scale/7) * q, same product as restore (CloverVector4.h:179-188)
            want = getattr(oracle, f"v{bits_}_restore")(ov, os_, n)[i]
        else:               # get(): (q * scale) / 127 (CloverVector8.h:136-139)
            want = np.float32(ov[i]) * os_[i >> 6] / np.float32(127.0)
        assert np.float32(q.get(i)) == want


@pytest.mark.parametrize("bits_", [4, 8])
def test_vector_zero_blocks(cb, oracle, bits_):
    V = cb.CloverVector4 if bits_ == 4 else cb.CloverVector8
    n = 384
    x = np.zeros(n, np.float32)
    x[64:128] = -0.0
    x[130], x[131], x[200] = 3.5, -0.0, -1e-30
    q = V(n)
    q.quantize(cb.CloverVector32(n, x))
    ov, os_ = getattr(oracle, f"v{bits_}_quantize")(x, n)
    assert np.array_equal(q.getData().cpu().numpy(), ov)
    assert np.array_equal(bits(q.getScales().cpu().numpy()), bits(os_))
    assert q.getScales()[0].item() == 1.0        # zero guard (CloverVector4.h:661-663)


@pytest.mark.parametrize("bits_", [4, 8])
@pytest.mark.parametrize("n", [64, 128, 1000, 4096, 65536 + 3, (1 << 20) + 128])
def test_vector_quantize_stochastic_matches_reference_stream(cb, oracle, n, bits_):
    """With an explicit key the device consumes the XORShift128+ stream exactly like the reference's
    sequential loop: packed bytes, scales AND the advanced key are identical."""
    V = cb.CloverVector4 if bits_ == 4 else cb.CloverVector8
    x = gen(oracle, n, "floats")
    st = oracle.xs_init(7, 9)
    ov, os_ = getattr(oracle, f"v{bits_}_quantize")(x, n, state=st)
    q = V(n)
    q.seed(7, 9)
    q.quantize(cb.CloverVector32(n, x))
    assert np.array_equal(q.getData().cpu().numpy(), ov)
    assert np.array_equal(bits(q.getScales().cpu().numpy()), bits(os_))
    assert np.array_equal(q.key, st), "PRNG key after quantize differs from the reference's"


@pytest.mark.parametrize("bits_", [4, 8])
@pytest.mark.parametrize("kind", ["floats", "ints"])
@pytest.mark.parametrize("n", [1, 128, 1000, 4096, 65536, 65536 + 1152])
def test_vector_dot_exact_order(cb, oracle, n, kind, bits_):
    """C1 (n=4096): the fp32 result is bit-identical to the reference SIMD dot (up to 65536 elements the CTA-parallel
    kernel, beyond - forced EXACT - the one-warp kernel)."""
    from clover_b200 import DOT_EXACT
    V = cb.CloverVector4 if bits_ == 4 else cb.CloverVector8
    x, y = gen(oracle, n, kind), gen(oracle, n, kind, skip=n + 17)
    qx, qy = V(n), V(n)
    qx.quantize(cb.CloverVector32(n, x))
    qy.quantize(cb.CloverVector32(n, y))
    xv, xs = getattr(oracle, f"v{bits_}_quantize")(x, n)
    yv, ys = getattr(oracle, f"v{bits_}_quantize")(y, n)
    want = getattr(oracle, f"v{bits_}_dot")(xv, xs, yv, ys, n)
    got = np.float32(qx.dot(qy, DOT_EXACT))
    assert got.view(np.uint32) == want.view(np.uint32), f"{float(got).hex()} vs {float(want).hex()}"
    if n <= 65536:
        assert np.float32(qx.dot(qy)).view(np.uint32) == want.view(np.uint32)   # AUTO picks EXACT here


def _dot_terms(xv, xs, yv, ys, n, bits_):
    """float64 evaluation of the reference's own per-block terms s_b * I_b, and sum |s_b * I_b|."""
    npad = n + (-n) % 128
    if bits_ == 4:
        def unpack(v):
            b = v.astype(np.int16)
            hi = b >> 4
            lo = ((b & 0xF) ^ 8) - 8
            return np.stack([hi, lo], 1).reshape(-1)
        a, b = unpack(xv), unpack(yv)
        s = (xs * np.float32(1.0 / 49.0)) * ys
    else:
        a, b = xv.astype(np.int32), yv.astype(np.int32)
        s = (xs * np.float32(1.0 / 127.0)) * (ys * np.float32(1.0 / 127.0))
    ib = (a.astype(np.int64) * b).reshape(npad // 64, 64).sum(1)
    t = s.astype(np.float64)[: npad // 64] * ib
    return t.sum(), np.abs(t).sum()


@pytest.mark.parametrize("bits_", [4, 8])
@pytest.mark.parametrize("n", [4096, 100000, 1 << 22])
def test_vector_dot_fast_mode(cb, oracle, n, bits_):
    """FAST mode: same exact integers, fp64 tree. Tolerance (SURVEY.md 8c rule 3):
    |gpu - ref| <= 4 * eps32 * sum_b |s_b * I_b|; and FAST is within 1 ulp of the fp64 evaluation."""
    from clover_b200 import DOT_FAST
    V = cb.CloverVector4 if bits_ == 4 else cb.CloverVector8
    x, y = gen(oracle, n, "floats"), gen(oracle, n, "floats", skip=n + 17)
    qx, qy = V(n), V(n)
    qx.quantize(cb.CloverVector32(n, x))
    qy.quantize(cb.CloverVector32(n, y))
    xv, xs = getattr(oracle, f"v{bits_}_quantize")(x, n)
    yv, ys = getattr(oracle, f"v{bits_}_quantize")(y, n)
    ref = float(getattr(oracle, f"v{bits_}_dot")(xv, xs, yv, ys, n))
    got = qx.dot(qy, DOT_FAST)
    exact64, sum_abs = _dot_terms(xv, xs, yv, ys, n, bits_)
    eps = float(np.finfo(np.float32).eps)
    assert abs(got - exact64) <= eps * abs(exact64) + 1e-30
    assert abs(got - ref) <= 4 * eps * sum_abs + abs(ref - exact64)
    assert got == qx.dot(qy, DOT_FAST), "FAST mode must be deterministic"


MAT_SHAPES = [(128, 128), (128, 256), (256, 128), (256, 384), (384, 640), (200, 300), (640, 1152), (1280, 1280),
              (128, 8192 + 128)]


@pytest.mark.parametrize("bits_", [4, 8])
@pytest.mark.parametrize("kind", ["floats", "ints"])
@pytest.mark.parametrize("shape", MAT_SHAPES)
def test_matrix_quantize_and_mvm(cb, oracle, shape, kind, bits_):
    """03_matrix.cpp:38-96 and :248-326 - packed matrix, its scales and the re-quantized mvm result."""
    from oracle.pyoracle import pad_matrix
    rows, cols = shape
    a = pad_matrix(gen(oracle, rows * cols, kind)[: rows * cols].reshape(rows, cols))
    R, Cc = a.shape
    xvec = gen(oracle, Cc, kind, skip=11)
    M = cb.CloverMatrix4 if bits_ == 4 else cb.CloverMatrix8
    V = cb.CloverVector4 if bits_ == 4 else cb.CloverVector8
    qa = M(rows, cols)
    qa.quantize(cb.CloverMatrix32(rows, cols, a))
    mv, ms = getattr(oracle, f"m{bits_}_quantize")(a)
    assert np.array_equal(qa.getData().cpu().numpy(), mv)
    assert np.array_equal(bits(qa.getScales().cpu().numpy()), bits(ms))
    assert qa.getBytes() == mv.nbytes + ms.nbytes

    qx, qy = V(Cc), V(R)
    qx.quantize(cb.CloverVector32(Cc, xvec))
    y32 = torch.zeros(R, dtype=torch.float32, device="cuda")
    qa.mvm(qx, qy, y32=y32)
    xv, xs = getattr(oracle, f"v{bits_}_quantize")(xvec, Cc)
    yv, ys, oy32 = getattr(oracle, f"m{bits_}_mvm")(mv, ms, R, Cc, xv, xs, want_f32=True)
    assert np.array_equal(bits(y32.cpu().numpy()), bits(oy32)), "fp32 row results (block_values)"
    assert np.array_equal(qy.getData().cpu().numpy(), yv)
    assert np.array_equal(bits(qy.getScales().cpu().numpy()), bits(ys))
    for (i, j) in ((0, 0), (R - 1, Cc - 1), (R // 2, Cc // 3)):
        q = mv.reshape(R, -1)
        if bits_ == 4:
            b = int(q[i, j // 2]); nib = (b >> 4) & 0xF if j % 2 == 0 else b & 0xF
            val = np.float32(ms[(i >> 6) * (Cc >> 6) + (j >> 6)] / np.float32(7.0)) * np.float32(nib - 16 if nib >= 8 else nib)
        else:
            val = np.float32(ms[(i >> 6) * (Cc >> 6) + (j >> 6)] / np.float32(127.0)) * np.float32(q[i, j])
        assert np.float32(qa.get(i, j)) == val


@pytest.mark.parametrize("shape", [(128, 128), (256, 384), (640, 1152), (128, 4224), (384, 2048)])
def test_matrix_mvm_f32(cb, oracle, shape):
    """mvm(V32,V32) (CloverMatrix4.h:1451-1547): bit-exact, far inside the reference's own 0.01 bound."""
    from oracle.pyoracle import pad_matrix
    rows, cols = shape
    a = pad_matrix(gen(oracle, rows * cols, "floats")[: rows * cols].reshape(rows, cols))
    xvec = gen(oracle, cols, "floats", skip=5)
    qa = cb.CloverMatrix4(rows, cols)
    qa.quantize(cb.CloverMatrix32(rows, cols, a))
    y = cb.CloverVector32(rows)
    qa.mvm(cb.CloverVector32(cols, xvec), y)
    mv, ms = oracle.m4_quantize(a)
    want = oracle.m4_mvm_f32(mv, ms, rows, cols, xvec)
    assert np.array_equal(bits(y.getData().cpu().numpy()[:rows]), bits(want))


@pytest.mark.parametrize("bits_", [4, 8])
@pytest.mark.parametrize("shape", [(128, 128), (256, 384), (384, 1152)])
def test_matrix_stochastic_matches_reference_stream(cb, oracle, shape, bits_):
    from oracle.pyoracle import pad_matrix
    rows, cols = shape
    a = pad_matrix(gen(oracle, rows * cols, "floats")[: rows * cols].reshape(rows, cols))
    xvec = gen(oracle, cols, "floats", skip=5)
    M = cb.CloverMatrix4 if bits_ == 4 else cb.CloverMatrix8
    V = cb.CloverVector4 if bits_ == 4 else cb.CloverVector8
    st = oracle.xs_init(123, 456)
    mv, ms = getattr(oracle, f"m{bits_}_quantize")(a, state=st)
    qa = M(rows, cols)
    qa.seed(123, 456)
    qa.quantize(cb.CloverMatrix32(rows, cols, a))
    assert np.array_equal(qa.getData().cpu().numpy(), mv)
    assert np.array_equal(bits(qa.getScales().cpu().numpy()), bits(ms))
    assert np.array_equal(qa.key, st)
    xv, xs = getattr(oracle, f"v{bits_}_quantize")(xvec, cols)
    yv, ys = getattr(oracle, f"m{bits_}_mvm")(mv, ms, rows, cols, xv, xs, state=st)
    qx, qy = V(cols), V(rows)
    qx.quantize(cb.CloverVector32(cols, xvec))
    qa.mvm(qx, qy)
    assert np.array_equal(qy.getData().cpu().numpy(), yv)
    assert np.array_equal(bits(qy.getScales().cpu().numpy()), bits(ys))
    assert np.array_equal(qa.key, st)


def _random_m8(cb, rows, cols, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    m = cb.CloverMatrix8(rows, cols)
    m.values.copy_(torch.randint(-127, 128, (rows * cols,), dtype=torch.int8, device="cuda", generator=g))
    m.scales.uniform_(0.05, 4.0, generator=g)
    return m


def _random_v8(cb, n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    v = cb.CloverVector8(n)
    v.values.copy_(torch.randint(-127, 128, (n,), dtype=torch.int8, device="cuda", generator=g))
    v.scales.uniform_(0.05, 4.0, generator=g)
    return v


@pytest.mark.parametrize("shape", [(4736 + 128, 128), (4736 + 128, 384), (148 * 32 * 3 + 128, 4096 + 128), (8192, 2048 * 17)])
def test_mvm4_pipelined_many_items_vs_oracle(cb, oracle, shape, monkeypatch):
    """CloverMatrix4::mvm through the 32-row-item TMA kernel with several items per CTA, half chunks at the row end
    (cols = 128 mod 256) and row blocks finished by two different CTAs: bit-for-bit against the oracle."""
    monkeypatch.setenv("CLOVER_GEMV_IMPL", "items32")       # read per call by the launcher
    rows, cols = shape
    A = _random_m4(cb, rows, cols, 61)
    g = torch.Generator(device="cuda").manual_seed(62)
    x = cb.CloverVector4(cols)
    from bench import random_nibbles
    x.values.copy_(random_nibbles(torch, cols // 2, g, torch.device("cuda")))
    x.scales.uniform_(0.05, 4.0, generator=g)
    y = cb.CloverVector4(rows)
    y32 = torch.zeros(rows, dtype=torch.float32, device="cuda")
    for _ in range(2):
        A.mvm(x, y, y32=y32)
    wv, ws, w32 = oracle.m4_mvm(A.values.cpu().numpy(), A.scales.cpu().numpy(), rows, cols,
                                x.values.cpu().numpy(), x.scales.cpu().numpy(), want_f32=True)
    assert np.array_equal(bits(y32.cpu().numpy()), bits(w32))
    assert np.array_equal(y.getData().cpu().numpy(), wv)
    assert np.array_equal(bits(y.getScales().cpu().numpy()[: rows // 64]), bits(ws[: rows // 64]))
    y2 = cb.CloverVector4(rows)
    A.mvm(x, y2)
    assert torch.equal(y2.values, y.values) and torch.equal(y2.scales, y.scales)


@pytest.mark.parametrize("impl", ["ring64", "items32", "simple"])
def test_mvm4_kernels_equal_default(impl):
    """Every GEMV kernel selectable with CLOVER_GEMV_IMPL gives the same bytes as the default choice (fresh processes)."""
    import os, subprocess, sys
    code = (
        "import torch, sys, hashlib; sys.path.insert(0, %r)\n"
        "from clover_b200 import containers as cb\n"
        "from bench import random_nibbles\n"
        "g = torch.Generator(device='cuda').manual_seed(3)\n"
        "h = hashlib.sha256()\n"
        "for (r, c) in [(128, 128), (640, 1152), (4864, 2176), (32768, 1024)]:\n"
        "    A = cb.CloverMatrix4(r, c); A.values.copy_(random_nibbles(torch, r * c // 2, g, torch.device('cuda'))); A.scales.uniform_(0.05, 4.0, generator=g)\n"
        "    x = cb.CloverVector4(c); x.values.copy_(random_nibbles(torch, c // 2, g, torch.device('cuda'))); x.scales.uniform_(0.05, 4.0, generator=g)\n"
        "    y = cb.CloverVector4(r); A.mvm(x, y); torch.cuda.synchronize()\n"
        "    h.update(y.values.cpu().numpy().tobytes()); h.update(y.scales.cpu().numpy().tobytes())\n"
        "print(h.hexdigest())\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    outs = []
    for env_impl in (None, impl):
        env = dict(os.environ)
        env.pop("CLOVER_GEMV_IMPL", None)
        if env_impl:
            env["CLOVER_GEMV_IMPL"] = env_impl
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, env=env)
        assert out.returncode == 0, out.stdout + out.stderr
        outs.append(out.stdout.strip().splitlines()[-1])
    assert outs[0] == outs[1]


@pytest.mark.parametrize("shape", [(4736 + 128, 256), (148 * 32 * 3 + 128, 2048 + 128), (8192, 1024 * 17)])
def test_mvm8_pipelined_many_items_vs_oracle(cb, oracle, shape):
    """CloverMatrix8::mvm through the persistent TMA-ring kernel when every CTA walks several 32-row work items, the
    ring wraps many times and the two halves of a row block are finished by different CTAs: fp32 row results, packed
    bytes and scales bit-for-bit against the oracle (inputs are random packed operands, not quantizer output)."""
    rows, cols = shape
    A, x = _random_m8(cb, rows, cols, 31), _random_v8(cb, cols, 32)
    y = cb.CloverVector8(rows)
    y32 = torch.zeros(rows, dtype=torch.float32, device="cuda")
    for _ in range(2):                          # second call: the block counters must have been re-armed
        A.mvm(x, y, y32=y32)
    wv, ws, w32 = oracle.m8_mvm(A.values.cpu().numpy(), A.scales.cpu().numpy(), rows, cols,
                                x.values.cpu().numpy(), x.scales.cpu().numpy(), want_f32=True)
    assert np.array_equal(bits(y32.cpu().numpy()), bits(w32))
    assert np.array_equal(y.getData().cpu().numpy(), wv)
    assert np.array_equal(bits(y.getScales().cpu().numpy()[: rows // 64]), bits(ws[: rows // 64]))
    y2 = cb.CloverVector8(rows)
    A.mvm(x, y2)                                # without a caller-provided fp32 buffer (internal scratch)
    assert torch.equal(y2.values, y.values) and torch.equal(y2.scales, y.scales)


def test_full_size_properties_c5(cb):
    """BASELINE config C5 (32768 x 32768, 8-bit): fp32 row results against an fp64 evaluation of the same exact block
    integers (rule-3 bound), and the re-quantized vector == the reference re-quantizer applied to those fp32 values
    (block absmax, 127/max as an IEEE divide, truncation) evaluated with torch on the same device."""
    n = 32768
    A, x = _random_m8(cb, n, n, 41), _random_v8(cb, n, 42)
    y = cb.CloverVector8(n)
    y32 = torch.zeros(n, dtype=torch.float32, device="cuda")
    A.mvm(x, y, y32=y32)
    torch.cuda.synchronize()
    hb = n // 64
    xs64 = x.values.view(hb, 64).to(torch.float16)
    prod = ((A.scales.view(hb, hb) * np.float32(1.0 / 127.0)) * (x.scales[:hb] * np.float32(1.0 / 127.0)).unsqueeze(0))
    ref = torch.empty(n, dtype=torch.float64, device="cuda")
    sum_abs = torch.empty(n, dtype=torch.float64, device="cuda")
    step = 2048
    for r0 in range(0, n, step):                # fp16 products of int8 values accumulate exactly in fp32 (<= 64 * 127^2)
        a = A.values.view(n, hb, 64)[r0:r0 + step].to(torch.float16)
        ib = torch.einsum("rbk,bk->rb", a.float(), xs64.float()).to(torch.float64)
        p = prod[r0 // 64:(r0 + step) // 64].repeat_interleave(64, dim=0).to(torch.float64)
        ref[r0:r0 + step] = (p * ib).sum(1)
        sum_abs[r0:r0 + step] = (p * ib).abs().sum(1)
    eps = float(np.finfo(np.float32).eps)
    assert bool(((y32.double() - ref).abs() <= 4 * eps * sum_abs + 1e-30).all())
    m = y32.abs().view(-1, 64).max(1).values
    m = torch.where(m == 0, torch.ones_like(m), m)
    scale = (torch.full_like(m, 127.0) / m).unsqueeze(1)
    q = torch.trunc(y32.abs().view(-1, 64) * scale) * torch.sign(y32.view(-1, 64))
    assert torch.equal(y.values.view(-1, 64).to(torch.float32), q)
    assert torch.equal(y.scales[: n // 64], m)


@pytest.mark.parametrize("bits_", [4, 8])
@pytest.mark.parametrize("shape", [(128, 128), (128, 256), (256, 128), (384, 640), (200, 300), (640, 1152), (2048 + 128, 4096 + 384)])
def test_matrix_transpose(cb, oracle, shape, bits_):
    """transpose (03_matrix.cpp:153-246 checks qm.get(i,j) == qt.get(j,i)): every packed byte and every scale of the
    transposed matrix against the oracle (itself pinned to all reference variants), and get() symmetry."""
    from oracle.pyoracle import pad_matrix
    rows, cols = shape
    a = pad_matrix(gen(oracle, rows * cols, "floats")[: rows * cols].reshape(rows, cols))
    R, Cc = a.shape
    M = cb.CloverMatrix4 if bits_ == 4 else cb.CloverMatrix8
    qa, qt = M(R, Cc), M(Cc, R)
    qa.quantize(cb.CloverMatrix32(R, Cc, a))
    qa.transpose(qt)
    tv, ts = getattr(oracle, f"m{bits_}_transpose")(qa.getData().cpu().numpy(), qa.getScales().cpu().numpy(), R, Cc)
    assert np.array_equal(qt.getData().cpu().numpy(), tv)
    assert np.array_equal(bits(qt.getScales().cpu().numpy()), bits(ts))
    for (i, j) in ((0, 0), (R - 1, Cc - 1), (R // 2, Cc // 3), (1, 2), (66, 127)):
        assert qa.get(i, j) == qt.get(j, i)
    with pytest.raises(cb.CloverSizeError):
        qa.transpose(M(R + 128, Cc))


@pytest.mark.parametrize("bits_", [4, 8])
def test_transpose_full_size_involution(cb, bits_):
    """16384 x 16384 (the GEMM operand size): transposing twice restores every byte and scale, and the transposed
    matrix really is the transpose (checked through an independent torch unpack on the device)."""
    n = 16384
    A = _random_m4(cb, n, n, 51) if bits_ == 4 else _random_m8(cb, n, n, 51)
    M = cb.CloverMatrix4 if bits_ == 4 else cb.CloverMatrix8
    T, B = M(n, n), M(n, n)
    A.transpose(T)
    T.transpose(B)
    assert torch.equal(A.values, B.values) and torch.equal(A.scales, B.scales)
    assert torch.equal(T.scales.view(n // 64, n // 64), A.scales.view(n // 64, n // 64).t())
    if bits_ == 8:
        assert torch.equal(T.values.view(n, n), A.values.view(n, n).t())
    else:
        def unpack(m):          # element 2i in the high nibble
            b = m.values.view(torch.uint8).view(n, n // 2)
            return torch.stack((b >> 4, b & 0xF), dim=2).view(n, n)
        assert torch.equal(unpack(T), unpack(A).t())


def test_mvm_size_mismatch_raises(cb):
    qa = cb.CloverMatrix4(128, 256)
    with pytest.raises(cb.CloverSizeError):
        qa.mvm(cb.CloverVector4(128), cb.CloverVector4(128))


@pytest.mark.parametrize("kind", ["floats", "ints"])
@pytest.mark.parametrize("shape", [(128, 128), (256, 384), (200, 300), (640, 1152), (1024, 8192 + 128)])
def test_matrix4_mvm_vector8(cb, oracle, shape, kind):
    """mixed precision CloverMatrix4::mvm(V8,V8) (CloverMatrix4.h:1093-1441): packed bytes, scales and the fp32
    row results bit-for-bit; with a key, the stochastic re-quantizer's stream position too."""
    from oracle.pyoracle import pad_matrix
    rows, cols = shape
    a = pad_matrix(gen(oracle, rows * cols, kind)[: rows * cols].reshape(rows, cols))
    R, Cc = a.shape
    xvec = gen(oracle, Cc, kind, skip=11)
    qa = cb.CloverMatrix4(R, Cc)
    qa.quantize(cb.CloverMatrix32(R, Cc, a))
    qx, qy = cb.CloverVector8(Cc), cb.CloverVector8(R)
    qx.quantize(cb.CloverVector32(Cc, xvec))
    y32 = torch.empty(R, dtype=torch.float32, device="cuda")
    qa.mvm(qx, qy, y32=y32)
    mv, ms = oracle.m4_quantize(a)
    xv, xs = oracle.v8_quantize(xvec, Cc)
    wv, ws, w32 = oracle.m4_mvm_v8(mv, ms, R, Cc, xv, xs, want_f32=True)
    assert np.array_equal(bits(y32.cpu().numpy()), bits(w32))
    assert np.array_equal(qy.getData().cpu().numpy(), wv)
    assert np.array_equal(bits(qy.getScales().cpu().numpy()[: R // 64]), bits(ws[: R // 64]))
    st = oracle.xs_init(5, 6)
    wv, ws = oracle.m4_mvm_v8(mv, ms, R, Cc, xv, xs, state=st)
    qa.seed(5, 6)
    qa.mvm(qx, qy)
    assert np.array_equal(qy.getData().cpu().numpy(), wv)
    assert np.array_equal(qa.key, st)


@pytest.mark.parametrize("bits_", [4, 8])
@pytest.mark.parametrize("a", [0.5, -1.75, 0.0])
@pytest.mark.parametrize("n", [1, 127, 128, 640, 1000, 65536 + 3])
def test_vector_scale_and_add(cb, oracle, n, a, bits_):
    """scaleAndAdd (02_vector.cpp:342-394): out-of-place and in-place, every byte and scale vs the oracle."""
    V = cb.CloverVector4 if bits_ == 4 else cb.CloverVector8
    x, y = gen(oracle, n, "floats"), gen(oracle, n, "wide", skip=321)
    qu, qv, qr = V(n), V(n), V(n)
    qu.quantize(cb.CloverVector32(n, x))
    qv.quantize(cb.CloverVector32(n, y))
    ou, osu = getattr(oracle, f"v{bits_}_quantize")(x, n)
    ov, osv = getattr(oracle, f"v{bits_}_quantize")(y, n)
    wr, wsr = oracle.scale_and_add(bits_, ou, osu, ov, osv, a, n)
    qu.scaleAndAdd(qv, a, qr)
    assert np.array_equal(qr.getData().cpu().numpy(), wr)
    assert np.array_equal(bits(qr.getScales().cpu().numpy()), bits(wsr))
    qu.scaleAndAdd(qv, a)                      # in place
    assert np.array_equal(qu.getData().cpu().numpy(), wr)
    assert np.array_equal(bits(qu.getScales().cpu().numpy()), bits(wsr))


@pytest.mark.parametrize("bits_", [4, 8])
def test_vector_scale_and_add_stochastic(cb, oracle, bits_):
    """with an explicit key the noise slots and the advanced key equal the sequential reference's"""
    V = cb.CloverVector4 if bits_ == 4 else cb.CloverVector8
    n = 4096 + 64
    x, y = gen(oracle, n, "floats"), gen(oracle, n, "ints", skip=77)
    qu, qv, qr = V(n), V(n), V(n)
    qu.quantize(cb.CloverVector32(n, x))
    qv.quantize(cb.CloverVector32(n, y))
    ou, osu = getattr(oracle, f"v{bits_}_quantize")(x, n)
    ov, osv = getattr(oracle, f"v{bits_}_quantize")(y, n)
    st = oracle.xs_init(11, 22)
    wr, wsr = oracle.scale_and_add(bits_, ou, osu, ov, osv, 0.5, n, st)
    qu.seed(11, 22)
    qu.scaleAndAdd(qv, 0.5, qr)
    assert np.array_equal(qr.getData().cpu().numpy(), wr)
    assert np.array_equal(bits(qr.getScales().cpu().numpy()), bits(wsr))
    assert np.array_equal(qu.key, st)


@pytest.mark.parametrize("mnk", [(128, 128, 128), (128, 256, 384), (256, 128, 1152)])
def test_gemm_vs_definition(cb, oracle, mnk):
    """GEMM extension: every C[i][j] vs the reference SIMD dot of the two row views (rule 3 tolerance)."""
    from oracle.pyoracle import pad_matrix
    M, N, K = mnk
    a = pad_matrix(gen(oracle, M * K, "floats")[: M * K].reshape(M, K))
    b = pad_matrix(gen(oracle, N * K, "floats", skip=999)[: N * K].reshape(N, K))
    qa, qb = cb.CloverMatrix4(M, K), cb.CloverMatrix4(N, K)
    qa.quantize(cb.CloverMatrix32(M, K, a))
    qb.quantize(cb.CloverMatrix32(N, K, b))
    c = qa.gemm(qb).cpu().numpy()
    av, as_ = oracle.m4_quantize(a)
    bv, bs = oracle.m4_quantize(b)
    want = oracle.m4_gemm(av, as_, bv, bs, K, 0, M, 0, N)
    # bound: 4 * eps * sum_kb |s_kb * I_kb| <= 4 * eps * K * max|a| * max|b| (crude but sufficient here)
    eps = float(np.finfo(np.float32).eps)
    bound = 4 * eps * K * float(np.abs(a).max()) * float(np.abs(b).max())
    assert np.max(np.abs(c.astype(np.float64) - want.astype(np.float64))) <= bound


def test_gemm_expand_e4m3_all_codes(cb):
    """nibble -> FP8 E4M3 expansion: every one of the 16 two's-complement codes, natural element order."""
    rows, cols = 128, 256
    g = torch.Generator(device="cuda").manual_seed(5)
    m = cb.CloverMatrix4(rows, cols)
    m.values.copy_(torch.randint(-128, 128, (rows * cols // 2,), dtype=torch.int8, device="cuda", generator=g))
    got = m.expand_e4m3().cpu().numpy()
    b = m.values.cpu().numpy().view(np.uint8)
    nib = np.empty(rows * cols, np.int32)
    nib[0::2], nib[1::2] = b >> 4, b & 0xF
    q = np.where(nib >= 8, nib - 16, nib)
    e4m3_mag = np.array([0x00, 0x38, 0x40, 0x44, 0x48, 0x4A, 0x4C, 0x4E, 0x50], np.uint8)
    want = e4m3_mag[np.abs(q)] | np.where(q < 0, 0x80, 0).astype(np.uint8)
    assert np.array_equal(got, want)


def _random_m4(cb, rows, cols, seed):
    from bench import random_nibbles
    g = torch.Generator(device="cuda").manual_seed(seed)
    m = cb.CloverMatrix4(rows, cols)
    m.values.copy_(random_nibbles(torch, rows * cols // 2, g, torch.device("cuda")))
    m.scales.uniform_(0.05, 4.0, generator=g)
    return m


@pytest.mark.parametrize("mnk", [(128, 128, 128), (128, 256, 128), (256, 384, 640), (1024, 1280, 2048),
                                 (2176, 1152, 1024), (4096, 4096, 4096)])
def test_gemm_tensor_core_equals_simt(cb, mnk):
    """tcgen05 kernel vs the DP4A kernel: exact integer slabs + the same sequential fp32 chain => identical bits.
    Shapes cover a single tile, the N tail (N = 128 mod 256), ragged tile counts and multi-wave persistence."""
    M, N, K = mnk
    A, B = _random_m4(cb, M, K, 11), _random_m4(cb, N, K, 12)
    c_tc = A.gemm(B, impl="tc")
    c_simt = A.gemm(B, impl="simt")
    torch.cuda.synchronize()
    assert torch.equal(c_tc.view(torch.int32), c_simt.view(torch.int32))
    # the split API gives the same result, and a padded leading dimension is honoured
    a8, b8 = A.expand_e4m3(), B.expand_e4m3()
    big = torch.full((M, N + 64), 7.0, device="cuda")
    A.gemm_expanded(a8, B, b8, out=big[:, :N])
    assert torch.equal(big[:, :N].view(torch.int32), c_simt.view(torch.int32))
    assert bool((big[:, N:] == 7.0).all())


def test_gemm_c4_full_size_properties(cb):
    """BASELINE config C4 (16384^3): tensor-core result == DP4A result on every element (both exact-integer slabs),
    and linearity in the scales: doubling sA doubles C bit-for-bit (power-of-two scaling is exact)."""
    n = 16384
    A, B = _random_m4(cb, n, n, 21), _random_m4(cb, n, n, 22)
    c_tc = A.gemm(B, impl="tc")
    c_simt = A.gemm(B, impl="simt")
    torch.cuda.synchronize()
    assert torch.equal(c_tc.view(torch.int32), c_simt.view(torch.int32))
    del c_simt
    A.scales.mul_(2.0)
    c2 = A.gemm(B, impl="tc")
    assert torch.equal(c2, c_tc * 2.0)


def test_full_size_properties_c2(cb, oracle):
    """BASELINE config C2 (n = 2^26): size-independent properties instead of a CPU re-computation.
    * |x - restore(quantize(x))| <= scale/7 per element (02_vector.cpp:181-221 'consistency')
    * every block's max element quantizes to 6 or 7, and no nibble is -8
    * dot(q, q) FAST equals the fp64 evaluation of sum s_b^2/49 * I_b from the packed bytes."""
    n = 1 << 26
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.rand(n, generator=g, device="cuda", dtype=torch.float32) * 2 - 1
    v32 = cb.CloverVector32(n)
    v32.values.copy_(x)
    q = cb.CloverVector4(n)
    q.quantize(v32)
    r = cb.CloverVector32(n)
    q.restore(r)
    scales = q.getScales()
    err = (r.values - x).abs().view(-1, 64)
    assert bool((err <= (scales / 7.0).unsqueeze(1) * (1 + 1e-6)).all())
    assert bool((scales == x.abs().view(-1, 64).max(1).values).all()), "scale must be the exact block absmax"
    b = q.getData().view(torch.uint8)
    hi, lo = (b >> 4), (b & 0xF)
    assert int(((hi == 8) | (lo == 8)).sum()) == 0, "nibble -8 must never be produced"
    from clover_b200 import DOT_FAST
    got = q.dot(q, DOT_FAST)
    sq = lambda t: (((t.to(torch.int16) ^ 8) - 8) ** 2).to(torch.int32)
    ib = (sq(hi) + sq(lo)).view(-1, 32).sum(1).to(torch.float64)
    s = ((scales * np.float32(1.0 / 49.0)) * scales).to(torch.float64)
    want = float((s * ib).sum())
    assert abs(got - want) <= float(np.finfo(np.float32).eps) * abs(want)


# ---------------------------------------------------------------------------------------------------------
# threshold: include/CloverVector4.h:1913-1973, include/CloverVector8.h:1680-1740;
# the reference's acceptance test is test/validate/02_vector.cpp:450-500 (sorted magnitudes of the survivors)
# ---------------------------------------------------------------------------------------------------------
def _abs_all(oracle, bits_, values, scales, n):
    """getAbs(i) for every i < n, vectorised with the oracle's arithmetic (fp32, one rounding per operation)"""
    v = np.asarray(values).view(np.int8)
    if bits_ == 4:
        b = v[: (n + 1) // 2].view(np.uint8).astype(np.int32)
        q = np.empty(2 * b.size, np.int32)
        q[0::2], q[1::2] = b >> 4, b & 15
        q = np.where(q >= 8, q - 16, q)[:n].astype(np.float32)
        s = (np.asarray(scales, np.float32) / np.float32(7.0))[np.arange(n) >> 6]
        return np.abs(s * q)
    q = v[:n].astype(np.float32)
    return np.abs((q * np.asarray(scales, np.float32)[np.arange(n) >> 6]) / np.float32(127.0))


THR_CASES = [(128, 64), (129, 1), (1000, 64), (1000, 999), (1000, 1000), (1000, 0), (2047, 64), (4096, 300), (8192 + 77, 2000),
             (65536, 4096)]


@pytest.mark.parametrize("bits_", [4, 8])
@pytest.mark.parametrize("kind", ["floats", "ints"])
@pytest.mark.parametrize("n,k", THR_CASES)
def test_vector_threshold_exact(cb, oracle, n, k, kind, bits_):
    """EXACT mode = the reference's sequential heap walk: every byte equals the oracle's, ties included
    ("ints" inputs in [-10, 10] produce many equal magnitudes)."""
    from clover_b200._lib import THRESHOLD_EXACT
    V = cb.CloverVector4 if bits_ == 4 else cb.CloverVector8
    x = gen(oracle, n, kind, skip=k)
    q = V(n)
    q.quantize(cb.CloverVector32(n, x))
    qv, qs = q.getData().cpu().numpy().copy(), q.getScales().cpu().numpy().copy()
    want = oracle.threshold(bits_, qv, qs, n, k)
    q.threshold(k, THRESHOLD_EXACT)
    assert np.array_equal(q.getData().cpu().numpy(), want)
    assert np.array_equal(bits(q.getScales().cpu().numpy()), bits(qs))          # scales are not touched


@pytest.mark.parametrize("bits_", [4, 8])
@pytest.mark.parametrize("kind", ["floats", "ints"])
@pytest.mark.parametrize("n,k", THR_CASES + [((1 << 20) + 77, 50000), ((1 << 22) + 77, 300001)])
def test_vector_threshold_fast(cb, oracle, n, k, kind, bits_):
    """FAST mode (radix select): exactly min(k, n) survivors with unchanged bits, every magnitude above the k-th largest
    kept, every one below cleared, ties by lowest index; its sorted magnitudes equal the oracle's (the reference's own
    acceptance criterion), and with no tie at the threshold the bytes equal the oracle's."""
    from clover_b200._lib import THRESHOLD_FAST
    V = cb.CloverVector4 if bits_ == 4 else cb.CloverVector8
    x = gen(oracle, n, kind, skip=k)
    q = V(n)
    q.quantize(cb.CloverVector32(n, x))
    qv, qs = q.getData().cpu().numpy().copy(), q.getScales().cpu().numpy().copy()
    mags = _abs_all(oracle, bits_, qv, qs, n)
    q.threshold(k, THRESHOLD_FAST)
    got = q.getData().cpu().numpy()
    kept = _abs_all(oracle, bits_, got, qs, n)
    if k >= n:
        assert np.array_equal(got, qv)
        return
    if k == 0:
        assert not got.any()
        return
    order = np.sort(mags)[::-1]
    t = order[k - 1]
    changed = kept != mags                                                      # cleared elements (non-zero before)
    assert np.all(kept[changed] == 0)                                           # an element is either untouched or cleared
    assert np.array_equal(kept[mags > t], mags[mags > t])                       # everything above the threshold survives
    assert not kept[mags < t].any()                                             # everything below is cleared
    ties = np.flatnonzero(mags == t)
    r = k - int((mags > t).sum())                                               # ties that survive: the lowest indices
    if t > 0:
        assert np.array_equal(kept[ties[:r]], mags[ties[:r]]) and not kept[ties[r:]].any()
        assert int((kept > 0).sum()) == k
    assert np.array_equal(np.sort(kept)[::-1][:k], order[:k])                   # 02_vector.cpp:450-500
    if n <= 70000:
        want = oracle.threshold(bits_, qv, qs, n, k)
        assert np.array_equal(np.sort(_abs_all(oracle, bits_, want, qs, n)), np.sort(kept))
        if len(ties) == r:                                                      # no tie-break needed: identical bytes
            assert np.array_equal(got, want)
    # pad nibbles / bytes stay zero
    pad_from = (n + 1) // 2 if bits_ == 4 else n
    assert not got[pad_from:].any()


@pytest.mark.parametrize("bits_", [4, 8])
@pytest.mark.parametrize("n,k", [(1000, 10), (32768, 777), (100000, 777), (300000, 150001), (3000000, 1234567)])
def test_vector_threshold_fast_all_equal(cb, oracle, n, k, bits_):
    """worst case for the tie rule and for histogram contention: ONE magnitude - the k lowest indices survive
    (single CTA up to 8192 elements, one 8-CTA cluster up to 262144, seven launches beyond)"""
    from clover_b200._lib import THRESHOLD_FAST
    V = cb.CloverVector4 if bits_ == 4 else cb.CloverVector8
    x = np.full(n, -0.75, np.float32)
    x[1::2] = 0.75
    q = V(n)
    q.quantize(cb.CloverVector32(n, x))
    before = q.getData().cpu().numpy().copy()
    q.threshold(k, THRESHOLD_FAST)
    got = q.getData().cpu().numpy()
    nb = k // 2 if bits_ == 4 else k                     # k is even in the 4-bit cases or handled below
    if bits_ == 4 and k % 2:
        assert np.array_equal(got[:nb], before[:nb])
        assert got.view(np.uint8)[nb] == (before.view(np.uint8)[nb] & 0xF0)     # element k-1 (even index) sits in the high nibble
        assert not got[nb + 1:].any()
    else:
        assert np.array_equal(got[:nb], before[:nb]) and not got[nb:].any()


@pytest.mark.parametrize("bits_", [4, 8])
def test_vector_threshold_auto_and_idempotent(cb, oracle, bits_):
    """AUTO = EXACT up to the limit, FAST beyond; thresholding twice with the same k changes nothing (size-independent
    property, checked at 2^22 elements)."""
    import clover_b200
    V = cb.CloverVector4 if bits_ == 4 else cb.CloverVector8
    limit = clover_b200.lib().clover_threshold_exact_limit()
    assert limit == 4096
    n, k = 4096, 100
    x = gen(oracle, n, "ints")
    q = V(n)
    q.quantize(cb.CloverVector32(n, x))
    want = oracle.threshold(bits_, q.getData().cpu().numpy().copy(), q.getScales().cpu().numpy().copy(), n, k)
    q.threshold(k)                                                              # AUTO
    assert np.array_equal(q.getData().cpu().numpy(), want)
    n, k = 1 << 22, 12345
    x = gen(oracle, n, "floats")
    q = V(n)
    q.quantize(cb.CloverVector32(n, x))
    q.threshold_parallel(k)
    once = q.getData().clone()
    assert int((_abs_all(oracle, bits_, once.cpu().numpy(), q.getScales().cpu().numpy(), n) > 0).sum()) == k
    q.threshold(k)
    assert torch.equal(q.getData(), once)


# ---------------------------------------------------------------------------------------------------------
# the reference's application loops (test/performance/01_measure.h:924-946 Q_IHT, :1000-1020 Q_GD): the whole
# iteration sequence - mvm, scaleAndAdd, mvm, scaleAndAdd, threshold - stays bit-identical to the oracle
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("bits_", [4, 8])
@pytest.mark.parametrize("algo", ["iht", "gd"])
def test_application_loops(cb, oracle, bits_, algo):
    from clover_b200 import apps
    from clover_b200._lib import THRESHOLD_EXACT
    M, N, K, iters, mu = 256, 512, 40, 4, np.float32(0.05)
    V = cb.CloverVector4 if bits_ == 4 else cb.CloverVector8
    Mx = cb.CloverMatrix4 if bits_ == 4 else cb.CloverMatrix8
    st = oracle.xs_init()
    phi32 = oracle.fill_floats(M * N, -1.0, 1.0, st).reshape(M, N) * np.float32(0.0625)
    y32 = oracle.fill_floats(M, -1.0, 1.0, st)
    Phi, PhiT = Mx(M, N), Mx(N, M)
    Phi.quantize(cb.CloverMatrix32(M, N, phi32))
    Phi.transpose(PhiT)
    y = V(M)
    y.quantize(cb.CloverVector32(M, y32))
    x, t1, t2, t3 = V(N), V(M), V(M), V(N)
    if algo == "iht":
        apps.Q_IHT(Phi, PhiT, x, y, t1, t2, t3, iters, K, float(mu), THRESHOLD_EXACT)
    else:
        apps.Q_GD(Phi, PhiT, x, y, t1, t2, t3, iters, float(mu))
    # the same loop on the oracle
    mq = getattr(oracle, f"m{bits_}_quantize")
    mvm = getattr(oracle, f"m{bits_}_mvm")
    tr = getattr(oracle, f"m{bits_}_transpose")
    pv, ps = mq(phi32)
    tv, ts = tr(pv, ps, M, N)
    yv, ys = getattr(oracle, f"v{bits_}_quantize")(y32, M)
    xv = np.zeros(N * bits_ // 8, np.int8)
    xs = np.ones(N // 64, np.float32)
    for _ in range(iters):
        t1v, t1s = mvm(pv, ps, M, N, xv, xs)
        t2v, t2s = oracle.scale_and_add(bits_, yv, ys, t1v, t1s, -1.0, M)
        t3v, t3s = mvm(tv, ts, N, M, t2v, t2s)
        xv, xs = oracle.scale_and_add(bits_, xv, xs, t3v, t3s, float(mu), N)
        if algo == "iht":
            xv = oracle.threshold(bits_, xv, xs, N, K)
    assert np.array_equal(x.getData().cpu().numpy(), xv)
    assert np.array_equal(bits(x.getScales().cpu().numpy()), bits(xs))
    assert x.getData().any()                                                    # the loop did something


def test_application_loop_cuda_graph(cb, oracle):
    """A whole Q_IHT call captured into a CUDA graph replays to the same bits as the eager call."""
    from clover_b200 import apps
    from clover_b200._lib import THRESHOLD_FAST
    M, N, K = 512, 2048, 100
    st = oracle.xs_init()
    phi32 = oracle.fill_floats(M * N, -1.0, 1.0, st).reshape(M, N) * np.float32(0.03)
    Phi, PhiT = cb.CloverMatrix4(M, N), cb.CloverMatrix4(N, M)
    Phi.quantize(cb.CloverMatrix32(M, N, phi32))
    Phi.transpose(PhiT)
    y = cb.CloverVector4(M)
    y.quantize(cb.CloverVector32(M, oracle.fill_floats(M, -1.0, 1.0, st)))
    x, t1, t2, t3 = cb.CloverVector4(N), cb.CloverVector4(M), cb.CloverVector4(M), cb.CloverVector4(N)
    run = lambda: apps.Q_IHT(Phi, PhiT, x, y, t1, t2, t3, 6, K, 0.05, THRESHOLD_FAST)
    run()
    torch.cuda.synchronize()
    want_v, want_s = x.getData().clone(), x.getScales().clone()
    assert int((want_v != 0).sum()) > 0
    graph = apps.capture(run)
    x.getData().fill_(0x55)                      # poison: the replay must rebuild x from scratch (Q_IHT starts with x.clear())
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(x.getData(), want_v) and torch.equal(x.getScales().view(torch.int32), want_s.view(torch.int32))
