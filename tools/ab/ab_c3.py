"""A/B of two builds of libclover_b200.so on the same box and in the same process (boxes differ by 5-10 %, and one box drifts by
as much while it warms up): the C3 mvm, a shard shape and the 8-bit mvm, order alternating between repetitions.

Build the baseline from an older commit next to this file, e.g.
    mkdir -p /tmp/ab && git archive <commit> clover_b200/csrc include | tar -x -C /tmp/ab && make -C /tmp/ab/clover_b200/csrc
    cp /tmp/ab/clover_b200/libclover_b200.so tools/ab/libclover_old.so
usage: python tools/ab/ab_c3.py
"""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bench import random_nibbles
HERE = os.path.dirname(os.path.abspath(__file__))
libs = {"old": C.CDLL(os.path.join(HERE, "libclover_old.so")), "new": C.CDLL(os.path.join(HERE, "..", "..", "clover_b200", "libclover_b200.so"))}
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
p = lambda t: C.c_void_p(t.data_ptr())
def run(fn, rows, cols, mbytes, xbytes, ybytes):
    vals = torch.empty(rows * cols * mbytes // 8, dtype=torch.int8, device=dev)
    vals.copy_(random_nibbles(torch, vals.numel(), g, dev))
    sc = torch.empty(rows * cols // 4096, dtype=torch.float32, device=dev).uniform_(0.25, 1.0, generator=g)
    xv = torch.empty(cols * xbytes // 8, dtype=torch.int8, device=dev); xv.copy_(random_nibbles(torch, xv.numel(), g, dev))
    xs = torch.empty(cols // 64, dtype=torch.float32, device=dev).uniform_(0.25, 1.0, generator=g)
    yv = torch.empty(rows * ybytes // 8, dtype=torch.int8, device=dev); ys = torch.empty(rows // 64, dtype=torch.float32, device=dev)
    for rep in range(6):
        for name, L in (list(libs.items()) if rep % 2 == 0 else list(libs.items())[::-1]):
            f = getattr(L, fn); f.restype = C.c_int
            call = lambda: f(p(vals), p(sc), C.c_uint64(rows), C.c_uint64(cols), p(xv), p(xs), p(yv), p(ys), None, None, None)
            for _ in range(5): assert call() == 0
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(100): call()
            e1.record(); e1.synchronize()
            print(fn, rows, cols, name, round(e0.elapsed_time(e1) * 10, 1), "us", flush=True)
run("clover_m4_mvm", 65536, 65536, 4, 4, 4)
run("clover_m4_mvm", 8192, 65536, 4, 4, 4)
run("clover_m8_mvm", 32768, 32768, 8, 8, 8)
