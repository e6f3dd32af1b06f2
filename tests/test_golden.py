"""Oracle vs the committed golden vectors (tests/golden/clover_golden.npz, generated from the unmodified
reference by oracle/gen_golden.py). CPU only. Also re-checks SURVEY.md 8c's known-answer table."""
import os

import numpy as np
import pytest

from oracle.pyoracle import fnv1a64, pad_matrix

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "clover_golden.npz")


@pytest.fixture(scope="module")
def g():
    return np.load(GOLDEN)


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


def test_prng_and_inputs(oracle, g):
    st = oracle.xs_init()
    assert np.array_equal(st, g["seed_state"])
    p = st.copy()
    for row in g["prng_first_16_calls"]:
        assert np.array_equal(oracle.xs_next(p), row)
    for name, n in (("a", 4096), ("b", 4096), ("c", 1000), ("d", 1000)):
        assert same_bits(oracle.fill_floats(n, -1.0, 1.0, st), g[name])
    M = oracle.fill_floats(256 * 384, -1.0, 1.0, st)
    assert same_bits(M.reshape(256, 384), g["M"])
    assert same_bits(oracle.fill_floats(384, -1.0, 1.0, st), g["v"])
    assert same_bits(oracle.fill_floats(384, -1.0, 1.0, st), g["w"])
    assert same_bits(oracle.fill_integers(1000, -10.0, 10.0, st), g["ints"])
    assert np.array_equal(st, g["state_after_inputs"])


def test_known_answer_table(g):
    # SURVEY.md 8c
    assert fnv1a64(g["a"].tobytes()) == "18dfa6095a9f4c3c"
    assert fnv1a64(g["v4_a_values"].tobytes()) == "f0c0b3dd721cc279"
    assert fnv1a64(g["v4_a_scales"].tobytes()) == "c42fc9d11429fbda"
    assert fnv1a64(g["v4_a_restore"].tobytes()) == "31e6456c16611b75"
    assert float(g["v4_dot_ab"][0]).hex() == "-0x1.ebbfb20000000p+3"            # C1
    assert fnv1a64(g["v4_c_values"].tobytes()) == "57dd345eb93197e7"
    assert float(g["v4_dot_cd"][0]).hex() == "0x1.47d7d00000000p+3"
    assert fnv1a64(g["v8_a_values"].tobytes()) == "e22c3448d27023f6"
    assert float(g["v8_dot_ab"][0]).hex() == "-0x1.4a8bf40000000p+4"
    assert fnv1a64(g["m4_mvm_values"].tobytes()) == "b204d807ebfa1302"
    assert fnv1a64(g["m4_mvm_f32"].tobytes()) == "71d6996935ac366d"
    assert fnv1a64(g["m8_mvm_values"].tobytes()) == "f52d4903d9bedb6c"


@pytest.mark.parametrize("bits", [4, 8])
def test_vectors(oracle, g, bits):
    for name, n in (("a", 4096), ("b", 4096), ("c", 1000), ("d", 1000), ("v", 384), ("ints", 1000)):
        v, s = getattr(oracle, f"v{bits}_quantize")(g[name], n)
        assert same_bits(v, g[f"v{bits}_{name}_values"]) and same_bits(s, g[f"v{bits}_{name}_scales"])
        assert same_bits(getattr(oracle, f"v{bits}_restore")(v, s, n), g[f"v{bits}_{name}_restore"])
    for p, q, n in (("a", "b", 4096), ("c", "d", 1000)):
        d = getattr(oracle, f"v{bits}_dot")(g[f"v{bits}_{p}_values"], g[f"v{bits}_{p}_scales"],
                                            g[f"v{bits}_{q}_values"], g[f"v{bits}_{q}_scales"], n)
        assert same_bits(np.array([d], np.float32), g[f"v{bits}_dot_{p}{q}"])


@pytest.mark.parametrize("bits", [4, 8])
def test_matrices(oracle, g, bits):
    M = pad_matrix(g["M"])
    mv, ms = getattr(oracle, f"m{bits}_quantize")(M)
    assert same_bits(mv, g[f"m{bits}_values"]) and same_bits(ms, g[f"m{bits}_scales"])
    yv, ys = getattr(oracle, f"m{bits}_mvm")(mv, ms, 256, 384, g[f"v{bits}_v_values"], g[f"v{bits}_v_scales"])
    assert same_bits(yv, g[f"m{bits}_mvm_values"]) and same_bits(ys, g[f"m{bits}_mvm_scales"])
    if bits == 4:
        assert same_bits(oracle.m4_mvm_f32(mv, ms, 256, 384, g["w"]), g["m4_mvm_f32"][:256])
        bv, bs = oracle.m4_quantize(pad_matrix(g["M"][:128].copy()))
        assert same_bits(oracle.m4_gemm(mv, ms, bv, bs, 384, 0, 256, 0, 128), g["m4_gemm_256x128"])


@pytest.mark.parametrize("bits", [4, 8])
def test_stochastic_with_key(oracle, g, bits):
    key = oracle.xs_init(7, 9)
    v, s = getattr(oracle, f"v{bits}_quantize")(g["c"], 1000, state=key)
    assert same_bits(v, g[f"sr_v{bits}_c_values"]) and same_bits(s, g[f"sr_v{bits}_c_scales"])
    assert np.array_equal(key, g[f"sr_v{bits}_key_after"])
    key = oracle.xs_init(123, 456)
    mv, ms = getattr(oracle, f"m{bits}_quantize")(pad_matrix(g["M"]), state=key)
    assert same_bits(mv, g[f"sr_m{bits}_values"]) and same_bits(ms, g[f"sr_m{bits}_scales"])
    assert np.array_equal(key, g[f"sr_m{bits}_key_after_quantize"])
    yv, ys = getattr(oracle, f"m{bits}_mvm")(mv, ms, 256, 384, g[f"v{bits}_v_values"], g[f"v{bits}_v_scales"], state=key)
    assert same_bits(yv, g[f"sr_m{bits}_mvm_values"]) and same_bits(ys, g[f"sr_m{bits}_mvm_scales"])
    assert np.array_equal(key, g[f"sr_m{bits}_key_after_mvm"])


# ---- scaleAndAdd, mixed mvm(V8), transpose, threshold (tests/golden/clover_golden_f.npz) ----------
GOLDEN_F = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "clover_golden_f.npz")


@pytest.fixture(scope="module")
def gf():
    return np.load(GOLDEN_F)


@pytest.mark.parametrize("bits", [4, 8])
def test_f_rows_vectors(oracle, g, gf, bits):
    for p, q, n in (("a", "b", 4096), ("c", "d", 1000)):
        u, su, v, sv = (g[f"v{bits}_{p}_values"], g[f"v{bits}_{p}_scales"], g[f"v{bits}_{q}_values"], g[f"v{bits}_{q}_scales"])
        for tag, alpha in (("p5", 0.5), ("m1", -1.0)):
            r, sr = oracle.scale_and_add(bits, u, su, v, sv, alpha, n)
            assert same_bits(r, gf[f"axpy{bits}_{p}{q}_{tag}_values"]) and same_bits(sr, gf[f"axpy{bits}_{p}{q}_{tag}_scales"])
    key = oracle.xs_init(11, 13)
    r, sr = oracle.scale_and_add(bits, g[f"v{bits}_c_values"], g[f"v{bits}_c_scales"], g[f"v{bits}_d_values"],
                                 g[f"v{bits}_d_scales"], 0.5, 1000, key)
    assert same_bits(r, gf[f"sr_axpy{bits}_cd_values"]) and same_bits(sr, gf[f"sr_axpy{bits}_cd_scales"])
    assert np.array_equal(key, gf[f"sr_axpy{bits}_key_after"])
    for name, n, k in (("ints", 1000, 64), ("a", 4096, 300), ("c", 1000, 999)):
        got = oracle.threshold(bits, g[f"v{bits}_{name}_values"], g[f"v{bits}_{name}_scales"], n, k)
        assert same_bits(got, gf[f"thr{bits}_{name}_k{k}"]), (name, k)


@pytest.mark.parametrize("bits", [4, 8])
def test_f_rows_transpose(oracle, g, gf, bits):
    tv, ts = getattr(oracle, f"m{bits}_transpose")(g[f"m{bits}_values"], g[f"m{bits}_scales"], 256, 384)
    assert same_bits(tv, gf[f"m{bits}_transpose_values"]) and same_bits(ts, gf[f"m{bits}_transpose_scales"])


def test_f_rows_mixed_mvm(oracle, g, gf):
    yv, ys = oracle.m4_mvm_v8(g["m4_values"], g["m4_scales"], 256, 384, g["v8_v_values"], g["v8_v_scales"])
    assert same_bits(yv, gf["m4_mvm_v8_values"]) and same_bits(ys, gf["m4_mvm_v8_scales"])
    key = oracle.xs_init(21, 22)
    yv, ys = oracle.m4_mvm_v8(g["m4_values"], g["m4_scales"], 256, 384, g["v8_v_values"], g["v8_v_scales"], state=key)
    assert same_bits(yv, gf["sr_m4_mvm_v8_values"]) and same_bits(ys, gf["sr_m4_mvm_v8_scales"])
    assert np.array_equal(key, gf["sr_m4_mvm_v8_key_after"])


def test_round2_rows_m8_mvm_f32_and_matrix_restore(oracle, g, gf):
    """CloverMatrix8::mvm(V32,V32) (CloverMatrix8.h:558-661) and the matrix restores (CloverMatrix4.h:266-301; 8-bit through
    get(i, j), CloverMatrix8.h:117-129): the oracle reproduces the bytes the compiled reference produced."""
    from oracle.pyoracle import fnv1a64
    assert same_bits(oracle.m8_mvm_f32(g["m8_values"], g["m8_scales"], 256, 384, g["w"]), gf["m8_mvm_f32"][:256])
    for bits in (4, 8):
        r = oracle.m_restore(bits, g[f"m{bits}_values"], g[f"m{bits}_scales"], 256, 384)
        assert same_bits(r[:8], gf[f"m{bits}_restore_rows8"])
        assert int(fnv1a64(r.tobytes()), 16) == int(gf[f"m{bits}_restore_fnv"][0])


def test_generators_known_answers(oracle, g):
    """SURVEY.md 8c known-answer table: the first draws of setRandomFloats(-1, 1) from the reference seeds, the matrix
    generator (the whole PADDED matrix is one run of the vector loop) and the integer variant."""
    st = oracle.xs_init()
    a = oracle.fill_floats(4096, -1.0, 1.0, st)
    assert [float(v).hex() for v in a[:4]] == ["0x1.ff90a00000000p-4", "-0x1.6047780000000p-3", "0x1.8b51800000000p-2", "0x1.3156700000000p-2"]
    assert same_bits(a[:4096], g["a"][:4096])
    st = g["seed_state"].copy()
    for name, n in (("a", 4096), ("b", 4096), ("c", 1000), ("d", 1000)):
        assert same_bits(oracle.fill_floats(n, -1.0, 1.0, st)[:n], g[name][:n])
    M = oracle.fill_floats(256 * 384, -1.0, 1.0, st)[: 256 * 384].reshape(256, 384)
    assert same_bits(M, g["M"])
    for name in ("v", "w"):
        assert same_bits(oracle.fill_floats(384, -1.0, 1.0, st)[:384], g[name][:384])
    assert same_bits(oracle.fill_integers(1000, -10.0, 10.0, st)[:1000], g["ints"][:1000])
    assert np.array_equal(st, g["state_after_inputs"])
