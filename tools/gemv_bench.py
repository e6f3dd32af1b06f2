"""Time the GEMV kernels (BASELINE C3 / C5) with CUDA events, operands resident in HBM.

usage: python tools/gemv_bench.py {4|8} [n] [reps=50]      (CLOVER_GEMV_IMPL=simple selects the plain-load kernels)
"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clover_b200 import containers as cb
from bench import random_nibbles, gemv_bytes, measured_peaks

def main():
    bits = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    n = int(sys.argv[2]) if len(sys.argv) > 2 else (65536 if bits == 4 else 32768)
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 50
    rows = int(sys.argv[4]) if len(sys.argv) > 4 else n          # optional: a row shard of the n x n matrix
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    M = (cb.CloverMatrix4 if bits == 4 else cb.CloverMatrix8)(rows, n)
    if bits == 4:
        M.values.copy_(random_nibbles(torch, rows * n // 2, g, dev))
    else:
        M.values.copy_(torch.randint(-127, 128, (rows * n,), dtype=torch.int8, device=dev, generator=g))
    M.scales.uniform_(0.25, 1.0, generator=g)
    V = cb.CloverVector4 if bits == 4 else cb.CloverVector8
    x, y = V(n), V(rows)
    v = cb.CloverVector32(n); v.values.uniform_(-1, 1, generator=g); x.quantize(v)
    for _ in range(5):
        M.mvm(x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        M.mvm(x, y)
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    by = gemv_bytes(rows, n, bits)
    peak = measured_peaks()[0]
    print(json.dumps({"bits": bits, "rows": rows, "n": n, "impl": os.environ.get("CLOVER_GEMV_IMPL", "tma"), "ms": ms,
                      "GBps": by / ms * 1e-6, "frac_hbm_peak": by / ms * 1e-6 / peak}))

if __name__ == "__main__":
    main()
