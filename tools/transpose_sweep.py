"""Time CloverMatrix4/8::transpose on a few shapes and check the result against torch's transpose of the unpacked matrix.

usage: python tools/transpose_sweep.py [reps=30]
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
    variants = ("default",)
    for bits in (4, 8):
        for rows, cols in ((16384, 16384), (8192, 32768), (32768, 8192), (4096, 4096 + 128)):
            M = (cb.CloverMatrix4 if bits == 4 else cb.CloverMatrix8)(rows, cols)
            if bits == 4:
                M.values.copy_(random_nibbles(torch, rows * cols // 2, g, dev))
            else:
                M.values.copy_(torch.randint(-127, 128, (rows * cols,), dtype=torch.int8, device=dev, generator=g))
            M.scales.uniform_(0.25, 1.0, generator=g)
            T = (cb.CloverMatrix4 if bits == 4 else cb.CloverMatrix8)(cols, rows)
            ref = None
            by = 2 * M.getBytes()
            for impl in variants:
                T.values.zero_(); T.scales.zero_()
                for _ in range(3):
                    M.transpose(T)
                torch.cuda.synchronize()
                ts = []
                for _ in range(3):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(reps):
                        M.transpose(T)
                    e1.record(); e1.synchronize()
                    ts.append(e0.elapsed_time(e1) / reps)
                ms = min(ts)
                if bits == 8:
                    good = torch.equal(T.values.view(cols, rows), M.values.view(rows, cols).t())
                else:
                    def unpack(m, r, c):
                        b = m.values.view(torch.uint8).view(r, c // 2)
                        return torch.stack((b >> 4, b & 0xF), dim=2).view(r, c)
                    good = torch.equal(unpack(T, cols, rows), unpack(M, rows, cols).t())
                good = good and torch.equal(T.scales.view(cols // 64, rows // 64), M.scales.view(rows // 64, cols // 64).t())
                same = "ok" if good else "WRONG"
                print(json.dumps({"bits": bits, "rows": rows, "cols": cols, "impl": impl, "us": round(ms * 1e3, 2),
                                  "GBps": round(by / ms * 1e-6, 1), "frac_hbm": round(by / ms * 1e-6 / peak, 3), "check": same}), flush=True)
            del M, T


if __name__ == "__main__":
    main()
