"""Time CloverMatrix4::mvm for row shards / IHT shapes under every kernel selection (CLOVER_GEMV_IMPL is read per call).

usage: python tools/gemv_shapes.py [reps=60] [bits=4]
"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clover_b200 import containers as cb
from bench import random_nibbles, gemv_bytes, measured_peaks


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    peak = measured_peaks()[0]
    bits = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    shapes4 = ((8192, 65536), (16384, 65536), (32768, 65536), (65536, 65536), (8192, 32768), (32768, 8192), (4096, 8192))
    shapes8 = ((32768, 32768), (16384, 32768), (8192, 32768), (4096, 32768), (32768, 8192), (8192, 16384), (16384, 4096))
    for rows, cols in (shapes4 if bits == 4 else shapes8):
        if bits == 4:
            M = cb.CloverMatrix4(rows, cols)
            M.values.copy_(random_nibbles(torch, rows * cols // 2, g, dev))
        else:
            M = cb.CloverMatrix8(rows, cols)
            M.values.copy_(torch.randint(-127, 128, (rows * cols,), dtype=torch.int8, device=dev, generator=g))
        M.scales.uniform_(0.25, 1.0, generator=g)
        V = cb.CloverVector4 if bits == 4 else cb.CloverVector8
        x, y = V(cols), V(rows)
        v = cb.CloverVector32(cols); v.values.uniform_(-1, 1, generator=g); x.quantize(v)
        ref = None
        for impl in (("ring64", "items32", "items32x2", "auto") if bits == 4 else ("items32", "items32x2", "auto")):
            if impl == "auto":
                os.environ.pop("CLOVER_GEMV_IMPL", None)
            else:
                os.environ["CLOVER_GEMV_IMPL"] = impl
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
            same = "ref" if ref is None else ("same" if torch.equal(ref[0], y.values) and torch.equal(ref[1], y.scales) else "DIFFERENT")
            if ref is None:
                ref = (y.values.clone(), y.scales.clone())
            by = gemv_bytes(rows, cols, bits)
            print(json.dumps({"rows": rows, "cols": cols, "impl": impl, "us": round(ms * 1e3, 2), "GBps": round(by / ms * 1e-6, 1),
                              "frac_hbm": round(by / ms * 1e-6 / peak, 3), "check": same}), flush=True)
        del M


if __name__ == "__main__":
    main()
