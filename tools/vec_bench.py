"""Time the vector-side kernels with CUDA events: dot (4/8-bit, FAST, rotating operand sets), threshold FAST, one IHT iteration.

usage: python tools/vec_bench.py [log2n=26]
"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clover_b200 import containers as cb, DOT_FAST, apps
from clover_b200._lib import THRESHOLD_FAST
from bench import cuda_time, measured_peaks


def main():
    n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 26)
    peak = measured_peaks()[0]
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(5)
    x32 = [cb.CloverVector32(n) for _ in range(2)]
    for v in x32:
        v.values.uniform_(-1.0, 1.0, generator=g)
    res = torch.empty(1, dtype=torch.float32, device=dev)
    out = {"n": n}
    for bits_, V in ((4, cb.CloverVector4), (8, cb.CloverVector8)):
        qs = [V(n) for _ in range(8)]
        for i, q in enumerate(qs):
            q.quantize(x32[i % 2])
        i = [0]
        def dot():
            k = i[0] % 4; qs[2 * k].dot_device(qs[2 * k + 1], res, DOT_FAST); i[0] += 1
        t = cuda_time(torch, dot, 80)
        b = 2 * qs[0].getBytes()
        out[f"dot{bits_}_fast"] = {"us": round(t * 1e6, 2), "GBps": round(b / t / 1e9, 1), "frac_hbm": round(b / t / 1e9 / peak, 3)}
        k = n // 64
        def thr():
            qs[i[0] % 8].threshold(k, THRESHOLD_FAST); i[0] += 1
        t = cuda_time(torch, thr, 16)
        b = 2 * qs[0].getBytes()          # the reference's own traffic model for threshold (01_measure.h:906)
        out[f"threshold{bits_}_fast_k=n/64"] = {"us": round(t * 1e6, 2), "GBps_ref_model": round(b / t / 1e9, 1)}
        del qs
    # thresholds at the IHT sizes (one CTA, one launch) and the EXACT heap walk
    from clover_b200._lib import THRESHOLD_EXACT
    for bits_, V in ((4, cb.CloverVector4), (8, cb.CloverVector8)):
        for nn in (4096, 32768, 131072):
            v32 = cb.CloverVector32(nn); v32.values.uniform_(-1.0, 1.0, generator=g)
            src, q = V(nn), V(nn)
            src.quantize(v32)
            def thr_small(mode):
                q.values.copy_(src.values); q.threshold(nn // 4, mode)
            t0 = cuda_time(torch, lambda: q.values.copy_(src.values), 50)
            out[f"threshold{bits_}_fast_n={nn}_k=n/4"] = {"us": round((cuda_time(torch, lambda: thr_small(THRESHOLD_FAST), 50) - t0) * 1e6, 2)}
            if nn <= 32768:
                out[f"threshold{bits_}_exact_n={nn}_k=n/4"] = {"us": round((cuda_time(torch, lambda: thr_small(THRESHOLD_EXACT), 5) - t0) * 1e6, 2)}
    # one IHT iteration at the reference's shape class (Phi M x N with N = 4M)
    M, N, K = 8192, 32768, 1024
    Phi, PhiT = cb.CloverMatrix4(M, N), cb.CloverMatrix4(N, M)
    Phi.values.copy_(torch.randint(-128, 128, (Phi.values.numel(),), dtype=torch.int8, device=dev, generator=g))
    Phi.scales.uniform_(0.01, 0.02, generator=g)
    Phi.transpose(PhiT)
    y = cb.CloverVector4(M)
    v = cb.CloverVector32(M); v.values.uniform_(-1, 1, generator=g); y.quantize(v)
    x, t1, t2, t3 = cb.CloverVector4(N), cb.CloverVector4(M), cb.CloverVector4(M), cb.CloverVector4(N)
    t = cuda_time(torch, lambda: apps.Q_IHT(Phi, PhiT, x, y, t1, t2, t3, 10, K, 0.01, THRESHOLD_FAST), 5)
    out["iht4_8192x32768_per_iteration"] = {"us": round(t * 1e5, 2), "matrix_bytes_per_iteration": 2 * Phi.getBytes(),
                                            "GBps": round(2 * Phi.getBytes() / (t / 10) / 1e9, 1)}
    graph = apps.capture(lambda: apps.Q_IHT(Phi, PhiT, x, y, t1, t2, t3, 10, K, 0.01, THRESHOLD_FAST))
    t = cuda_time(torch, graph.replay, 5)
    out["iht4_8192x32768_per_iteration_cuda_graph"] = {"us": round(t * 1e5, 2), "GBps": round(2 * Phi.getBytes() / (t / 10) / 1e9, 1)}
    # the five steps of one iteration on their own (each captured in a graph of 20 repeats: device time without host gaps)
    steps = {"mvm_Phi_8192x32768": lambda: Phi.mvm(x, t1), "scaleAndAdd_M": lambda: y.scaleAndAdd(t1, -1.0, t2),
             "mvm_PhiT_32768x8192": lambda: PhiT.mvm(t2, t3), "scaleAndAdd_N": lambda: x.scaleAndAdd(t3, 0.01),
             "threshold_N": lambda: x.threshold(K, THRESHOLD_FAST)}
    for name, fn in steps.items():
        gr = apps.capture(lambda: [fn() for _ in range(20)])
        out["iht_step_" + name] = {"us": round(cuda_time(torch, gr.replay, 5) / 20 * 1e6, 2)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
