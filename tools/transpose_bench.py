"""Time CloverMatrix4/8::transpose with CUDA events. usage: python tools/transpose_bench.py {4|8} [n=16384] [reps=50]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clover_b200 import containers as cb
from bench import measured_peaks

def main():
    bits = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 50
    g = torch.Generator(device="cuda").manual_seed(1)
    M = cb.CloverMatrix4 if bits == 4 else cb.CloverMatrix8
    A, T = M(n, n), M(n, n)
    A.values.copy_(torch.randint(-128, 128, (A.values.numel(),), dtype=torch.int8, device="cuda", generator=g))
    A.scales.uniform_(0.25, 1.0, generator=g)
    for _ in range(5):
        A.transpose(T)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        A.transpose(T)
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    by = 2 * A.getBytes()                       # every byte read once and written once
    print(json.dumps({"bits": bits, "n": n, "ms": ms, "GBps": by / ms * 1e-6, "frac_hbm_peak": by / ms * 1e-6 / measured_peaks()[0]}))

if __name__ == "__main__":
    main()
