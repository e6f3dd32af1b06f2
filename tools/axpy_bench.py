"""Time scaleAndAdd at n = 2^26 for both widths."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clover_b200 import containers as cb
from bench import cuda_time
n = 1 << 26
g = torch.Generator(device="cuda").manual_seed(5)
x32 = [cb.CloverVector32(n) for _ in range(2)]
for v in x32: v.values.uniform_(-1.0, 1.0, generator=g)
for bits_, V in ((4, cb.CloverVector4), (8, cb.CloverVector8)):
    qs = [V(n) for _ in range(6)]
    for i, q in enumerate(qs): q.quantize(x32[i % 2])
    i = [0]
    def axpy():
        k = i[0] % 2; qs[3 * k].scaleAndAdd(qs[3 * k + 1], 0.5, qs[3 * k + 2]); i[0] += 1
    t = min(cuda_time(torch, axpy, 20) for _ in range(3))
    b = 3 * qs[0].getBytes()
    print(bits_, "bit", round(t * 1e6, 1), "us", round(b / t / 1e9), "GB/s", flush=True)
    del qs
