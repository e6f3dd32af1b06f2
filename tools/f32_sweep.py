"""Time mvm(V32,V32) for both matrix widths under either work-item height (CLOVER_GEMV_IMPL=rows16 / rows8 is read per
call) and compare the bits with the plain-load kernel.

usage: python tools/f32_sweep.py [reps=30]
"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clover_b200 import containers as cb
from bench import random_nibbles, measured_peaks


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    peak = measured_peaks()[0]
    for bits in (4, 8):
        for rows, cols in ((32768, 32768), (8192, 32768), (32768, 8192), (65536, 16384), (4096, 4096 + 128)):
            M = (cb.CloverMatrix4 if bits == 4 else cb.CloverMatrix8)(rows, cols)
            if bits == 4:
                M.values.copy_(random_nibbles(torch, rows * cols // 2, g, dev))
            else:
                M.values.copy_(torch.randint(-127, 128, (rows * cols,), dtype=torch.int8, device=dev, generator=g))
            M.scales.uniform_(0.25, 1.0, generator=g)
            x, y = cb.CloverVector32(cols), cb.CloverVector32(rows)
            x.values.uniform_(-1, 1, generator=g)
            os.environ["CLOVER_GEMV_IMPL"] = "simple"
            M.mvm(x, y)
            torch.cuda.synchronize()
            ref = y.values.clone()
            by = M.getBytes() + 4 * (rows + cols)
            for impl in ("rows16", "rows8", "auto"):
                if impl == "auto":
                    os.environ.pop("CLOVER_GEMV_IMPL", None)
                else:
                    os.environ["CLOVER_GEMV_IMPL"] = impl
                y.values.zero_()
                for _ in range(3):
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
                same = "same" if torch.equal(ref.view(torch.int32), y.values.view(torch.int32)) else "DIFFERENT"
                print(json.dumps({"bits": bits, "rows": rows, "cols": cols, "impl": impl, "us": round(ms * 1e3, 2),
                                  "GBps": round(by / ms * 1e-6, 1), "frac_hbm": round(by / ms * 1e-6 / peak, 3), "check": same}), flush=True)
            del M


if __name__ == "__main__":
    main()
