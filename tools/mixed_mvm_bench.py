"""Time the mixed-precision mvm (4-bit matrix x CloverVector8) under every kernel selection; checks the bits agree."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clover_b200 import containers as cb
from bench import cuda_time, random_nibbles, measured_peaks
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(5)
peak = measured_peaks()[0]
for rows, cols in ((32768, 32768), (8192, 32768), (32768, 8192), (4096, 16384)):
    m4 = cb.CloverMatrix4(rows, cols)
    m4.values.copy_(random_nibbles(torch, rows * cols // 2, g, dev)); m4.scales.uniform_(0.25, 1.0, generator=g)
    x8, y8 = cb.CloverVector8(cols), cb.CloverVector8(rows)
    v = cb.CloverVector32(cols); v.values.uniform_(-1, 1, generator=g); x8.quantize(v)
    ref = None
    for impl in ("simple", "items32", "items32x2", None):
        if impl is None: os.environ.pop("CLOVER_GEMV_IMPL", None)
        else: os.environ["CLOVER_GEMV_IMPL"] = impl
        t = min(cuda_time(torch, lambda: m4.mvm(x8, y8), 30) for _ in range(3))
        same = "ref" if ref is None else ("same" if torch.equal(ref[0], y8.values) and torch.equal(ref[1], y8.scales) else "DIFFERENT")
        if ref is None: ref = (y8.values.clone(), y8.scales.clone())
        b = m4.getBytes() + x8.getBytes() + y8.getBytes()
        print(rows, cols, impl or "auto", round(t * 1e6, 1), "us", round(b / t / 1e9), "GB/s", round(b / t / 1e9 / peak, 3), same, flush=True)
    del m4
