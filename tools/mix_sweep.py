"""Time the mixed-precision CloverMatrix4::mvm(V8,V8) under either ring geometry (CLOVER_GEMV_IMPL is read per call) and
compare the bytes with the plain-load kernel (CLOVER_GEMV_IMPL=simple), whose arithmetic the parity tests pin to the oracle.

usage: python tools/mix_sweep.py [reps=40]
"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clover_b200 import containers as cb
from bench import random_nibbles, measured_peaks


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    peak = measured_peaks()[0]
    variants = ("ring4", "ring8", "auto")
    for rows, cols in ((32768, 32768), (8192, 32768), (32768, 8192), (4096, 32768), (16384, 16384), (1024, 1152)):
        M = cb.CloverMatrix4(rows, cols)
        M.values.copy_(random_nibbles(torch, rows * cols // 2, g, dev))
        M.scales.uniform_(0.25, 1.0, generator=g)
        x, y = cb.CloverVector8(cols), cb.CloverVector8(rows)
        v = cb.CloverVector32(cols); v.values.uniform_(-1, 1, generator=g); x.quantize(v)
        os.environ["CLOVER_GEMV_IMPL"] = "simple"
        M.mvm(x, y)
        torch.cuda.synchronize()
        ref = (y.values.clone(), y.scales.clone())
        by = M.getBytes() + x.getBytes() + y.getBytes()
        for impl in variants:
            if impl == "auto":
                os.environ.pop("CLOVER_GEMV_IMPL", None)
            else:
                os.environ["CLOVER_GEMV_IMPL"] = impl
            y.values.zero_(); y.scales.zero_()
            for _ in range(5):
                M.mvm(x, y)
            torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    M.mvm(x, y)
                e1.record(); e1.synchronize()
                ts.append(e0.elapsed_time(e1) / reps)
            ms = min(ts)
            same = "same" if torch.equal(ref[0], y.values) and torch.equal(ref[1], y.scales) else "DIFFERENT"
            print(json.dumps({"rows": rows, "cols": cols, "impl": impl, "us": round(ms * 1e3, 2), "GBps": round(by / ms * 1e-6, 1),
                              "frac_hbm": round(by / ms * 1e-6 / peak, 3), "check": same}), flush=True)
        del M


if __name__ == "__main__":
    main()
