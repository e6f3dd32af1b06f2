"""GPU parity for the SURVEY 8 pieces added in round 2 (VERDICT r01 missing #1, #2, #4 and weak #3):
CloverMatrix8::mvm(V32,V32), the re-designed CloverMatrix4::mvm(V32,V32) ring kernel at awkward shapes, matrix restore,
the reference's input generators on the device. Everything bit-for-bit against the oracle. Needs a B200."""
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


def _quantized(cb, oracle, bits_, rows, cols, kind, seed_skip=0):
    """a quantized matrix built by the product from reference-generated fp32 input + the oracle's packed twin"""
    from oracle.pyoracle import pad_matrix
    st = oracle.xs_init()
    oracle.xs_skip(st, seed_skip)
    fill = oracle.fill_integers if kind == "ints" else oracle.fill_floats
    lo, hi = (-10.0, 10.0) if kind == "ints" else (-1.0, 1.0)
    a = pad_matrix(fill(rows * cols, lo, hi, st)[: rows * cols].reshape(rows, cols))
    x = fill(a.shape[1], lo, hi, st)
    M = cb.CloverMatrix4 if bits_ == 4 else cb.CloverMatrix8
    qa = M(rows, cols)
    qa.quantize(cb.CloverMatrix32(rows, cols, a))
    mv, ms = getattr(oracle, f"m{bits_}_quantize")(a)
    return qa, mv, ms, x, a.shape


F32_SHAPES = [(128, 128), (256, 384), (640, 1152), (128, 4224), (384, 2048), (200, 300), (4864, 2176), (1024, 16384 + 128),
              (148 * 8 * 32 + 256, 640)]


@pytest.mark.parametrize("impl", [None, "rows8", "rows16"])
@pytest.mark.parametrize("bits_", [4, 8])
@pytest.mark.parametrize("kind", ["floats", "ints"])
@pytest.mark.parametrize("shape", F32_SHAPES)
def test_matrix_mvm_f32_ring_vs_oracle(cb, oracle, shape, kind, bits_, impl, monkeypatch):
    """mvm(V32,V32) (CloverMatrix4.h:1451-1547, CloverMatrix8.h:558-661) through the per-warp TMA-ring kernels (default
    dispatch; 8-row items with two chunks per stage and 16-row items forced): partial chunks and partial stages at the row
    end (cols = 64 / 128 / 192 mod 256, odd chunk counts), more work items than warps (several items per warp, ring wrap
    across items), one block per row - every fp32 result bit-for-bit."""
    if impl is None:
        monkeypatch.delenv("CLOVER_GEMV_IMPL", raising=False)
    else:
        monkeypatch.setenv("CLOVER_GEMV_IMPL", impl)          # read per call by the launcher
    qa, mv, ms, x, (R, Cc) = _quantized(cb, oracle, bits_, *shape, kind)
    y = cb.CloverVector32(R)
    qa.mvm(cb.CloverVector32(Cc, x), y)
    want = getattr(oracle, f"m{bits_}_mvm_f32")(mv, ms, R, Cc, x)
    assert np.array_equal(bits(y.getData().cpu().numpy()[:R]), bits(want))
    # the plain-load kernel (unaligned operands) gives the same bits
    monkeypatch.setenv("CLOVER_GEMV_IMPL", "simple")
    y2 = cb.CloverVector32(R)
    qa.mvm(cb.CloverVector32(Cc, x), y2)
    assert torch.equal(y2.values.view(torch.int32), y.values.view(torch.int32))


@pytest.mark.parametrize("bits_", [4, 8])
def test_matrix_mvm_f32_bench_shape_properties(cb, oracle, bits_):
    """The shape bench.py times (32768^2): 2048 sampled rows against the oracle (the oracle walks only those rows)."""
    from bench import random_nibbles
    n = 32768
    g = torch.Generator(device="cuda").manual_seed(5 + bits_)
    M = cb.CloverMatrix4 if bits_ == 4 else cb.CloverMatrix8
    A = M(n, n)
    if bits_ == 4:
        A.values.copy_(random_nibbles(torch, n * n // 2, g, torch.device("cuda")))
    else:
        A.values.copy_(torch.randint(-127, 128, (n * n,), dtype=torch.int8, device="cuda", generator=g))
    A.scales.uniform_(0.05, 4.0, generator=g)
    x = cb.CloverVector32(n)
    x.values.uniform_(-1, 1, generator=g)
    y = cb.CloverVector32(n)
    A.mvm(x, y)
    got = y.values.cpu().numpy()
    xh = x.values.cpu().numpy()
    per_row = n * bits_ // 8
    for r0 in (0, 64 * 37, 64 * 300 + 64, n - 1024):          # 4 x 512 rows (whole 64-row tiles: the scale rows line up)
        rows = 512
        mv = A.values[r0 * per_row:(r0 + rows) * per_row].cpu().numpy()
        ms = A.scales[(r0 // 64) * (n // 64):((r0 + rows) // 64) * (n // 64)].cpu().numpy()
        want = getattr(oracle, f"m{bits_}_mvm_f32")(mv, ms, rows, n, xh)
        assert np.array_equal(bits(got[r0:r0 + rows]), bits(want)), r0


@pytest.mark.parametrize("bits_", [4, 8])
@pytest.mark.parametrize("shape", [(128, 128), (256, 384), (200, 300), (1152, 2048 + 128)])
def test_matrix_restore(cb, oracle, shape, bits_):
    qa, mv, ms, _, (R, Cc) = _quantized(cb, oracle, bits_, *shape, "floats")
    out = cb.CloverMatrix32(R, Cc)
    qa.restore(out)
    want = oracle.m_restore(bits_, mv, ms, R, Cc)
    assert np.array_equal(bits(out.getData().cpu().numpy()), bits(want))
    for (i, j) in ((0, 0), (R - 1, Cc - 1), (R // 2, Cc // 3)):      # restore(i, j) == get(i, j), the reference's definition for 8 bits
        assert np.float32(qa.get(i, j)) == want[i, j]


@pytest.mark.parametrize("n", [1, 7, 8, 9, 127, 128, 1000, 4096, 65536 + 3, (1 << 22) + 5])
def test_generators_match_reference_stream(cb, oracle, n):
    """setRandomFloats / setRandomInteger (CloverVector32.h:712-783) on the device: values, the untouched pad and the
    advanced key pair equal the reference's, left-over elements (n % 8, one call each) included."""
    for integer in (False, True):
        st = oracle.xs_init(11, 12)
        want = (oracle.fill_integers if integer else oracle.fill_floats)(n, -10.0 if integer else -1.0, 10.0 if integer else 1.0, st)
        v = cb.CloverVector32(n)
        key = oracle.xs_init(11, 12)
        if integer:
            v.setRandomInteger(-10.0, 10.0, key)
        else:
            v.setRandomFloats(-1.0, 1.0, key)
        assert np.array_equal(bits(v.getData().cpu().numpy()), bits(want)), integer
        assert np.array_equal(key, st), "key after the generator call"


def test_generators_known_answers_and_matrix(cb, oracle):
    """SURVEY.md 8c known answers from the reference seeds (test/random/00_random.cpp:42), then a CloverMatrix32 and a
    second vector drawn from the SAME advancing key pair, as the reference's harnesses do."""
    key = oracle.xs_init()
    st = oracle.xs_init()
    a = cb.CloverVector32(4096)
    a.setRandomFloats(-1.0, 1.0, key)
    assert [float(v).hex() for v in a.getData()[:4].cpu().numpy()] == ["0x1.ff90a00000000p-4", "-0x1.6047780000000p-3", "0x1.8b51800000000p-2", "0x1.3156700000000p-2"]
    oracle.fill_floats(4096, -1.0, 1.0, st)
    M = cb.CloverMatrix32(200, 300)                     # padded to 256 x 384: the generator fills the pad as well (CloverMatrix32.h:294)
    M.setRandomFloats(-1.0, 1.0, key)
    want = oracle.fill_floats(256 * 384, -1.0, 1.0, st)[: 256 * 384].reshape(256, 384)
    assert np.array_equal(bits(M.getData().cpu().numpy()), bits(want))
    w = cb.CloverVector32(384)
    w.setRandomInteger(-10.0, 10.0, key)
    assert np.array_equal(bits(w.getData().cpu().numpy()), bits(oracle.fill_integers(384, -10.0, 10.0, st)))
    assert np.array_equal(key, st)
    # a container without an explicit key seeds its own (the reference: RDRAND) and stays in range
    z = cb.CloverVector32(1000)
    z.setRandomFloats(-2.0, 3.0)
    zv = z.getData()[:1000]
    assert float(zv.min()) >= -2.0 and float(zv.max()) <= 3.0 and float(zv.std()) > 0.5
