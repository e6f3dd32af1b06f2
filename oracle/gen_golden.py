"""Generate tests/golden/clover_golden.npz from the UNMODIFIED reference (oracle/_ref).

Run where /root/reference exists:  python oracle/gen_golden.py
Inputs follow SURVEY.md 8c: keys from avx_xorshift128plus_init(445560390295639063, 2935984234003016713)
(test/random/00_random.cpp:42), then drawn sequentially from that one key pair with
setRandomFloats(-1, 1): a[4096], b[4096], c[1000], d[1000], M[256x384], v[384], w[384].
Outputs are stored raw (packed bytes, fp32 scales, fp32 results as bit patterns) so that the oracle, and
through it the CUDA path, is pinned to bytes the real reference produced.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Reference, build, fnv1a64, pad_matrix  # noqa: E402


def main():
    build(ref=True)
    ref, ref_sr = Reference(False), Reference(True)
    st = ref.xs_init()
    g = {"seed_state": st.copy()}
    prng = ref.xs_init()
    g["prng_first_16_calls"] = np.stack([ref.xs_next(prng) for _ in range(16)])
    a = ref.fill_floats(4096, -1, 1, st)
    b = ref.fill_floats(4096, -1, 1, st)
    c = ref.fill_floats(1000, -1, 1, st)
    d = ref.fill_floats(1000, -1, 1, st)
    m32 = Reference._M(ref, 32, 256, 384)
    ref.lib.ref_m32_fill_floats(m32.h, C.c_float(-1.0), C.c_float(1.0), st.ctypes.data_as(C.c_void_p))
    M = m32.values.copy().reshape(256, 384)
    v = ref.fill_floats(384, -1, 1, st)
    w = ref.fill_floats(384, -1, 1, st)
    ints = ref.fill_integers(1000, -10, 10, st)
    g.update(a=a, b=b, c=c, d=d, M=M, v=v, w=w, ints=ints, state_after_inputs=st.copy())

    for name, x, n in (("a", a, 4096), ("b", b, 4096), ("c", c, 1000), ("d", d, 1000), ("v", v, 384), ("ints", ints, 1000)):
        for bits in (4, 8):
            qv, qs = getattr(ref, f"v{bits}_quantize")(x, n)
            g[f"v{bits}_{name}_values"], g[f"v{bits}_{name}_scales"] = qv, qs
            g[f"v{bits}_{name}_restore"] = getattr(ref, f"v{bits}_restore")(qv, qs, n)
    for bits in (4, 8):
        for (p, q_, n) in (("a", "b", 4096), ("c", "d", 1000)):
            g[f"v{bits}_dot_{p}{q_}"] = np.array([getattr(ref, f"v{bits}_dot")(
                g[f"v{bits}_{p}_values"], g[f"v{bits}_{p}_scales"], g[f"v{bits}_{q_}_values"], g[f"v{bits}_{q_}_scales"], n)],
                np.float32)
        mv, ms, h = getattr(ref, f"m{bits}_quantize")(pad_matrix(M))
        g[f"m{bits}_values"], g[f"m{bits}_scales"] = mv, ms
        yv, ys = getattr(ref, f"m{bits}_mvm")(h, g[f"v{bits}_v_values"], g[f"v{bits}_v_scales"])
        g[f"m{bits}_mvm_values"], g[f"m{bits}_mvm_scales"] = yv, ys
        if bits == 4:
            g["m4_mvm_f32"] = ref.m4_mvm_f32(h, w).copy()
            bt, bts, hb = ref.m4_quantize(pad_matrix(M[:128].copy()))
            g["m4_gemm_256x128"] = ref.m4_gemm(h, hb, 0, 256, 0, 128)
    # stochastic rounding with an explicit key (the reference built WITHOUT the disable flag)
    for bits in (4, 8):
        key = ref_sr.xs_init(7, 9)
        qv, qs = getattr(ref_sr, f"v{bits}_quantize")(c, 1000, state=key)
        g[f"sr_v{bits}_c_values"], g[f"sr_v{bits}_c_scales"], g[f"sr_v{bits}_key_after"] = qv, qs, key.copy()
        key = ref_sr.xs_init(123, 456)
        mv, ms, h = getattr(ref_sr, f"m{bits}_quantize")(pad_matrix(M), state=key)
        g[f"sr_m{bits}_values"], g[f"sr_m{bits}_scales"], g[f"sr_m{bits}_key_after_quantize"] = mv, ms, key.copy()
        xv, xs = getattr(ref, f"v{bits}_quantize")(v, 384)
        yv, ys = getattr(ref_sr, f"m{bits}_mvm")(h, xv, xs, state=key)
        g[f"sr_m{bits}_mvm_values"], g[f"sr_m{bits}_mvm_scales"], g[f"sr_m{bits}_key_after_mvm"] = yv, ys, key.copy()

    # cross-check against the known-answer table of SURVEY.md 8c before writing anything
    assert fnv1a64(a.tobytes()[: 4096 * 4]) == "18dfa6095a9f4c3c"
    assert fnv1a64(g["v4_a_values"].tobytes()) == "f0c0b3dd721cc279"
    assert fnv1a64(g["v4_a_scales"].tobytes()) == "c42fc9d11429fbda"
    assert float(g["v4_dot_ab"][0]).hex() == "-0x1.ebbfb20000000p+3"
    assert float(g["v8_dot_ab"][0]).hex() == "-0x1.4a8bf40000000p+4"
    assert fnv1a64(g["m4_mvm_values"].tobytes()) == "b204d807ebfa1302"
    out = os.path.join(ROOT, "tests", "golden", "clover_golden.npz")
    np.savez_compressed(out, **{k: np.ascontiguousarray(val) for k, val in g.items()})
    print("wrote", out, os.path.getsize(out), "bytes,", len(g), "arrays")
    golden_f(ref, ref_sr, g)


def golden_f(ref, ref_sr, g):
    """SURVEY.md 8f rows - scaleAndAdd, mixed mvm(V8), transpose, threshold - on the same inputs -> clover_golden_f.npz"""
    f = {}
    for bits in (4, 8):
        for (p, q_, n) in (("a", "b", 4096), ("c", "d", 1000)):
            u, su, v, sv = (g[f"v{bits}_{p}_values"], g[f"v{bits}_{p}_scales"], g[f"v{bits}_{q_}_values"], g[f"v{bits}_{q_}_scales"])
            for tag, alpha in (("p5", 0.5), ("m1", -1.0)):
                r, sr = ref.scale_and_add(bits, u, su, v, sv, alpha, n)
                f[f"axpy{bits}_{p}{q_}_{tag}_values"], f[f"axpy{bits}_{p}{q_}_{tag}_scales"] = r, sr
        key = ref_sr.xs_init(11, 13)
        r, sr = ref_sr.scale_and_add(bits, g[f"v{bits}_c_values"], g[f"v{bits}_c_scales"], g[f"v{bits}_d_values"],
                                     g[f"v{bits}_d_scales"], 0.5, 1000, state=key)
        f[f"sr_axpy{bits}_cd_values"], f[f"sr_axpy{bits}_cd_scales"], f[f"sr_axpy{bits}_key_after"] = r, sr, key.copy()
        h = getattr(ref, f"m{bits}_from")(g[f"m{bits}_values"], g[f"m{bits}_scales"], 256, 384)
        tv, ts = getattr(ref, f"m{bits}_transpose")(h)
        f[f"m{bits}_transpose_values"], f[f"m{bits}_transpose_scales"] = tv, ts
        for name, n, k in (("ints", 1000, 64), ("a", 4096, 300), ("c", 1000, 999)):
            f[f"thr{bits}_{name}_k{k}"] = ref.threshold(bits, g[f"v{bits}_{name}_values"], g[f"v{bits}_{name}_scales"], n, k)
    h4 = ref.m4_from(g["m4_values"], g["m4_scales"], 256, 384)
    yv, ys = ref.m4_mvm_v8(h4, g["v8_v_values"], g["v8_v_scales"])
    f["m4_mvm_v8_values"], f["m4_mvm_v8_scales"] = yv, ys
    key = ref_sr.xs_init(21, 22)
    h4s = ref_sr.m4_from(g["m4_values"], g["m4_scales"], 256, 384)
    yv, ys = ref_sr.m4_mvm_v8(h4s, g["v8_v_values"], g["v8_v_scales"], state=key)
    f["sr_m4_mvm_v8_values"], f["sr_m4_mvm_v8_scales"], f["sr_m4_mvm_v8_key_after"] = yv, ys, key.copy()
    # round 2: CloverMatrix8::mvm(V32,V32) (CloverMatrix8.h:558-661) and the matrix restores (CloverMatrix4.h:266-301;
    # 8-bit through get(i, j), CloverMatrix8.h:117-129) on the same M, w
    h8 = ref.m8_from(g["m8_values"], g["m8_scales"], 256, 384)
    f["m8_mvm_f32"] = ref.m8_mvm_f32(h8, g["w"]).copy()
    for bits, h in ((4, h4), (8, h8)):
        r = ref.m_restore(bits, h)
        f[f"m{bits}_restore_rows8"] = r[:8].copy()
        f[f"m{bits}_restore_fnv"] = np.array([int(fnv1a64(r.tobytes()), 16)], np.uint64)
    out = os.path.join(ROOT, "tests", "golden", "clover_golden_f.npz")
    np.savez_compressed(out, **{k: np.ascontiguousarray(val) for k, val in f.items()})
    print("wrote", out, os.path.getsize(out), "bytes,", len(f), "arrays")


if __name__ == "__main__":
    main()
