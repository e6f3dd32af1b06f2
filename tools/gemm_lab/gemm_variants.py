"""Walk through every 4-bit GEMM kernel variant / measurement mode in ONE process and time each with CUDA events.

usage: python tools/gemm_variants.py [n=16384] [reps=7] [spec ...]
  spec = KERNEL[:PROBE[:EPI]] e.g. "tc", "tc:2", "p192::1", "v4::3" ("-" = unset); default: a built-in list.
Prints one line per variant: median / best kernel ms, POP/s, SM clock sampled while 12 launches are in flight.
Measurement modes (PROBE != 0, EPI 3/4) compute wrong results on purpose; they only time parts of the pipeline.
"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clover_b200 import containers as cb
from bench import random_nibbles

DEFAULT = ["tc", "tc:1", "tc:2", "split", "p192::0", "p192::1", "p192::2", "p192::3", "p192::4", "pair", "pipe1"]


def sm_clock():
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        return pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
    except Exception:
        return -1


def main():
    args = sys.argv[1:]
    n = int(args[0]) if args else 16384
    reps = int(args[1]) if len(args) > 1 else 7
    specs = args[2:] or DEFAULT
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    A, B = cb.CloverMatrix4(n, n), cb.CloverMatrix4(n, n)
    for m in (A, B):
        m.values.copy_(random_nibbles(torch, n * n // 2, g, dev)); m.scales.uniform_(0.25, 1.0, generator=g)
    out = torch.empty(n, n, device=dev)
    a8, b8 = A.expand_e4m3(), B.expand_e4m3()
    ops = 2.0 * n ** 3
    ref = None
    for spec in specs:
        parts = (spec.split(":") + ["", ""])[:3]
        for key, val in zip(("CLOVER_GEMM_KERNEL", "CLOVER_GEMM_PROBE", "CLOVER_GEMM_EPI"), parts):
            if val in ("", "-"):
                os.environ.pop(key, None)
            else:
                os.environ[key] = val
        fn = lambda: A.gemm_expanded(a8, B, b8, out=out)
        try:
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            # sustained: 12 launches in flight, clock sampled while they run
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(12):
                fn()
            e1.record()
            time.sleep(ts[len(ts) // 2] * 6e-3)
            clk = sm_clock()
            e1.synchronize()
            sus = e0.elapsed_time(e1) / 12
            res = {"spec": spec, "n": n, "ms_med": round(ts[len(ts) // 2], 4), "ms_best": round(ts[0], 4), "ms_sustained": round(sus, 4),
                   "POPS_best": round(ops / ts[0] * 1e-12, 3), "POPS_sustained": round(ops / sus * 1e-12, 3), "sm_mhz_under_load": clk}
            if parts[1] in ("", "-", "0") and parts[2] in ("", "-", "0", "1", "2"):      # a real (non-probe) variant: compare the bits
                if ref is None:
                    ref = out.clone()
                    res["check"] = "reference"
                else:
                    res["check"] = "bit-identical" if torch.equal(out, ref) else "MISMATCH max|d|=%g" % (out - ref).abs().max().item()
        except Exception as e:                                                          # keep going: one broken variant must not hide the others
            res = {"spec": spec, "error": str(e)[:200]}
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
