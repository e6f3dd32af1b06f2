"""Time the 4-bit GEMM (BASELINE C4) pieces with CUDA events: expansion pass, tcgen05 kernel, whole call.

usage: python tools/gemm_bench.py [n=16384] [reps=10]
"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clover_b200 import containers as cb
from bench import random_nibbles

INT8_PEAK_TOPS = 4600.0     # tools/mma_probe peak, kind::i8, measured on this pool (profiles/r01_mma_probe.txt)

def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    A, B = cb.CloverMatrix4(n, n), cb.CloverMatrix4(n, n)
    for m in (A, B):
        m.values.copy_(random_nibbles(torch, n * n // 2, g, dev)); m.scales.uniform_(0.25, 1.0, generator=g)
    out = torch.empty(n, n, device=dev)
    a8, b8 = A.expand_e4m3(), B.expand_e4m3()
    ops = 2.0 * n ** 3
    res = {"n": n}
    med, best = timed(lambda: (A.expand_e4m3(out=a8), B.expand_e4m3(out=b8)), reps)
    res["expand_ms"] = med
    res["expand_GBps"] = (n * n / 2 + n * n) * 2 / med * 1e-6
    med, best = timed(lambda: A.gemm_expanded(a8, B, b8, out=out), reps)
    res["kernel_ms"], res["kernel_best_ms"] = med, best
    res["kernel_TOPS"] = ops / med * 1e-9
    res["kernel_frac_int8_peak"] = res["kernel_TOPS"] / INT8_PEAK_TOPS
    med, best = timed(lambda: A.gemm(B, out=out), reps)
    res["total_ms"] = med
    med2, best2 = timed(lambda: A.gemm_expanded(a8, B, b8, out=out), reps)
    res["kernel_ms_again"] = med2
    res["total_TOPS"] = ops / med * 1e-9
    res["total_frac_int8_peak"] = res["total_TOPS"] / INT8_PEAK_TOPS
    print(json.dumps(res))

if __name__ == "__main__":
    main()
