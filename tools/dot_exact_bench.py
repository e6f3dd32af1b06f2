"""Time the exact-order dot (bit-identical to the reference's SIMD dot) at the AUTO sizes. usage: python tools/dot_exact_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clover_b200 import containers as cb, DOT_EXACT
from bench import cuda_time
g = torch.Generator(device="cuda").manual_seed(1)
res = torch.empty(1, device="cuda")
for bits in (4, 8):
    for n in (4096, 16384, 65536):
        V = cb.CloverVector4 if bits == 4 else cb.CloverVector8
        a, b = V(n), V(n)
        v = cb.CloverVector32(n); v.values.uniform_(-1, 1, generator=g); a.quantize(v)
        v.values.uniform_(-1, 1, generator=g); b.quantize(v)
        t = cuda_time(torch, lambda: a.dot_device(b, res, DOT_EXACT), 200)
        print(f"dot{bits} exact n={n}: {t * 1e6:.2f} us per call", flush=True)
