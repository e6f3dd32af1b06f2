"""Launch ONE hot-path kernel a few times at its BASELINE.json size - the command ncu wraps.

usage: python tools/profile_driver.py {gemv4|gemv8|quantize4|quantize8|dot4|mquantize4|gemm4|quantize4_sr|transpose4|transpose8|
                                        threshold4_cluster|threshold8_cluster|threshold4_large|iht} [iters]
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clover_b200 import containers as cb, DOT_FAST
from bench import random_nibbles

what = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)

if what in ("gemv4", "gemv8"):
    bits = 4 if what == "gemv4" else 8
    n = 65536 if bits == 4 else 32768
    M = (cb.CloverMatrix4 if bits == 4 else cb.CloverMatrix8)(n, n)
    if bits == 4:
        M.values.copy_(random_nibbles(torch, n * n // 2, g, dev))
    else:
        M.values.copy_(torch.randint(-127, 128, (n * n,), dtype=torch.int8, device=dev, generator=g))
    M.scales.uniform_(0.25, 1.0, generator=g)
    V = cb.CloverVector4 if bits == 4 else cb.CloverVector8
    x, y = V(n), V(n)
    v = cb.CloverVector32(n); v.values.uniform_(-1, 1, generator=g); x.quantize(v)
    fn = lambda: M.mvm(x, y)
elif what in ("quantize4", "quantize8", "quantize4_sr"):
    n = 1 << 26
    v = cb.CloverVector32(n); v.values.uniform_(-1, 1, generator=g)
    q = (cb.CloverVector8 if what == "quantize8" else cb.CloverVector4)(n)
    if what.endswith("_sr"):
        q.seed(7, 9)
    fn = lambda: q.quantize(v)
elif what == "dot4":
    n = 1 << 26
    v = cb.CloverVector32(n)
    qs = []
    for _ in range(8):
        v.values.uniform_(-1, 1, generator=g); q = cb.CloverVector4(n); q.quantize(v); qs.append(q)
    res = torch.empty(1, device=dev)
    k = [0]
    def fn():
        i = k[0] % 4; qs[2 * i].dot_device(qs[2 * i + 1], res, DOT_FAST); k[0] += 1
elif what == "mquantize4":
    n = 16384
    a = cb.CloverMatrix32(n, n); a.values.uniform_(-1, 1, generator=g)
    q = cb.CloverMatrix4(n, n)
    fn = lambda: q.quantize(a)
elif what in ("transpose4", "transpose8"):
    n = 16384
    M = cb.CloverMatrix4 if what == "transpose4" else cb.CloverMatrix8
    A, T = M(n, n), M(n, n)
    A.values.copy_(torch.randint(-128, 128, (A.values.numel(),), dtype=torch.int8, device=dev, generator=g))
    A.scales.uniform_(0.25, 1.0, generator=g)
    fn = lambda: A.transpose(T)
elif what == "gemm4":
    n = int(os.environ.get("GEMM_N", "16384"))
    A, B = cb.CloverMatrix4(n, n), cb.CloverMatrix4(n, n)
    for m in (A, B):
        m.values.copy_(random_nibbles(torch, n * n // 2, g, dev)); m.scales.uniform_(0.25, 1.0, generator=g)
    out = torch.empty(n, n, device=dev)
    fn = lambda: A.gemm(B, out=out)
elif what in ("threshold4_cluster", "threshold4_large", "threshold8_cluster"):
    from clover_b200 import THRESHOLD_FAST
    n = 32768 if what.endswith("cluster") else 1 << 26
    V = cb.CloverVector8 if what.startswith("threshold8") else cb.CloverVector4
    v = cb.CloverVector32(n); v.values.uniform_(-1, 1, generator=g)
    src, q = V(n), V(n)
    src.quantize(v)
    def fn():
        q.values.copy_(src.values); q.threshold(n // 4 if n <= 32768 else n // 64, THRESHOLD_FAST)
elif what == "mvm4v8":
    n = 32768
    M = cb.CloverMatrix4(n, n)
    M.values.copy_(random_nibbles(torch, n * n // 2, g, dev)); M.scales.uniform_(0.25, 1.0, generator=g)
    x, y = cb.CloverVector8(n), cb.CloverVector8(n)
    v = cb.CloverVector32(n); v.values.uniform_(-1, 1, generator=g); x.quantize(v)
    fn = lambda: M.mvm(x, y)
elif what in ("mvmf32", "mvm8f32"):
    n = 32768
    if what == "mvmf32":
        M = cb.CloverMatrix4(n, n)
        M.values.copy_(random_nibbles(torch, n * n // 2, g, dev))
    else:
        M = cb.CloverMatrix8(n, n)
        M.values.copy_(torch.randint(-127, 128, (n * n,), dtype=torch.int8, device=dev, generator=g))
    M.scales.uniform_(0.25, 1.0, generator=g)
    x, y = cb.CloverVector32(n), cb.CloverVector32(n)
    x.values.uniform_(-1, 1, generator=g)
    fn = lambda: M.mvm(x, y)
elif what in ("axpy4", "axpy8"):
    n = 1 << 26
    V = cb.CloverVector4 if what == "axpy4" else cb.CloverVector8
    v = cb.CloverVector32(n)
    qs = []
    for _ in range(6):
        v.values.uniform_(-1, 1, generator=g); q = V(n); q.quantize(v); qs.append(q)
    k = [0]
    def fn():
        i = k[0] % 2; qs[3 * i].scaleAndAdd(qs[3 * i + 1], 0.5, qs[3 * i + 2]); k[0] += 1
elif what == "iht":
    from clover_b200 import THRESHOLD_FAST, apps
    M, N, K = 8192, 32768, 1024
    Phi, PhiT = cb.CloverMatrix4(M, N), cb.CloverMatrix4(N, M)
    Phi.values.copy_(torch.randint(-128, 128, (Phi.values.numel(),), dtype=torch.int8, device=dev, generator=g))
    Phi.scales.uniform_(0.01, 0.02, generator=g)
    Phi.transpose(PhiT)
    y = cb.CloverVector4(M)
    v = cb.CloverVector32(M); v.values.uniform_(-1, 1, generator=g); y.quantize(v)
    x, t1, t2, t3 = cb.CloverVector4(N), cb.CloverVector4(M), cb.CloverVector4(M), cb.CloverVector4(N)
    fn = lambda: apps.Q_IHT(Phi, PhiT, x, y, t1, t2, t3, 2, K, 0.01, THRESHOLD_FAST)
else:
    raise SystemExit(__doc__)

torch.cuda.synchronize()
for _ in range(iters):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    fn()
e1.record(); torch.cuda.synchronize()
print(what, "avg ms", e0.elapsed_time(e1) / iters)
