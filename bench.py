#!/usr/bin/env python
"""bench.py - the headline benchmark of clover_b200 (contract: see the task statement / DESIGN.md §Measurement).

Workload (BASELINE.json configs[2], "C3"): CloverMatrix4::mvm, 65536 x 65536 4-bit matrix times a 65536-element
CloverVector4, result re-quantized to a CloverVector4 (include/CloverMatrix4.h:777-1083), stochastic rounding
disabled (the parity configuration). One "step" = one mvm over the whole matrix. The 2 GiB matrix is far larger
than the 126 MB L2, so every step streams it from HBM ("inputs larger than L2").

  value      = algorithmic bytes moved per second, all GPUs, inputs resident in HBM
               (reference bytes model qv + qa + qr, test/performance/01_measure.h:717)
  e2e        = same metric through the container API with the per-step operands in HOST memory: the product
               vector is copied from pinned host memory and the result vector is read back every step; the
               matrix stays resident in HBM like the reference's matrix object stays in RAM between calls.
  roofline   = the GEMV kernel against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline = the reference's own mvm_parallel (oracle/_ref, compiled from /root/reference) on this host
  extras     = the other BASELINE.json configs (C1, C2, C4, C5), each timed with CUDA events

N > 1 (torchrun, one rank per GPU): rows sharded in 64-row blocks, total work fixed -> "scaling": "strong". The exchange
step is, by default, FUSED into the GEMV kernel: its epilogue stores each re-quantized 64-row block into every peer's
result vector over NVLink and synchronises with flags (no NCCL call, one kernel per step and rank);
`--exchange allreduce` is north_star's single NCCL allreduce of the fp32 output, `--exchange allgather` its cheaper twin.

`--impl reference` times the reference CPU implementation of the same path instead (rank 0 only).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS = COLS = 65536
METRIC = "int4_gemv_effective_GBps"
INT8_PEAK_TOPS = 4600.0      # measured on this pool with tools/mma_probe (kind::i8, burst 4586-4607, sustained 4602)


def gemv_bytes(rows, cols, bits=4):
    """qv.getBytes() + qa.getBytes() + qr.getBytes() (test/performance/01_measure.h:717)."""
    per = {4: 0.5, 8: 1.0}[bits]
    vec = lambda n: int(n * per) + (n // 64) * 4
    return vec(cols) + int(rows * cols * per) + (rows // 64) * (cols // 64) * 4 + vec(rows)


def ncu_traffic(key):
    """DRAM bytes per launch of a kernel/config from the committed `ncu --set full` captures (profiles/ncu_traffic.json); None if not captured"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(key, {}).get("bytes")
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region. In-process NVML polled by a background thread about
    every millisecond (nvidia-smi -lms cannot resolve a timed region of a few milliseconds: VERDICT r01 weak #9), every
    sample time-stamped; mark() records the start / end of the timed region and summary() keeps the samples in between.
    If the region was shorter than one poll, the nearest sample on either side is reported and flagged."""
    REASONS = (("hw_slowdown", "nvmlClocksThrottleReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
               ("sw_thermal_slowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksThrottleReasonSwPowerCap"))

    def __init__(self, index):
        self.index, self.samples, self.marks, self.stop, self.t, self.nv, self.h, self.max_mhz = index, [], [], False, None, None, None, None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def _poll(self):
        nv, h = self.nv, self.h
        while not self.stop:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((time.perf_counter(), int(mhz), int(r)))
            except Exception:
                pass
            time.sleep(0.0005)

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_mhz = int(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            t0 = time.time()
            while not self.samples and time.time() - t0 < 2.0:     # wait for the first sample
                time.sleep(0.001)
        except Exception:
            self.nv = None
        return self

    def mark(self):
        self.marks.append(time.perf_counter())

    def __exit__(self, *a):
        self.stop = True
        if self.t:
            self.t.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0, "source": "nvml unavailable"}
        lo, hi = (self.marks[0], self.marks[-1]) if len(self.marks) >= 2 else (self.samples[0][0], self.samples[-1][0])
        inside = [x for x in self.samples if lo <= x[0] <= hi]
        nearest = False
        if not inside:       # region shorter than one poll: the closest sample before and after it
            mid = 0.5 * (lo + hi)
            inside = sorted(self.samples, key=lambda x: abs(x[0] - mid))[:2]
            nearest = True
        reasons = set()
        for _, _, r in inside:
            for name, const in self.REASONS:
                if r & getattr(self.nv, const, 0):
                    reasons.add(name)
        out = {"sm_mhz": int(statistics.median(x[1] for x in inside)), "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons),
               "samples": len(inside), "source": "nvml, in-process, during the timed region", "window_ms": (hi - lo) * 1e3}
        if nearest:
            out["source"] = "nvml, in-process: the timed region was shorter than one poll, nearest samples around it"
        return out


# ----------------------------------------------------------------------------------------------------------------
# synthetic operands
# ----------------------------------------------------------------------------------------------------------------
def random_nibbles(torch, nbytes, gen, device):
    """Packed 4-bit values uniform in [-7, 7] (the quantizer never produces -8)."""
    out = torch.empty(nbytes, dtype=torch.int8, device=device)
    step = 1 << 28
    for o in range(0, nbytes, step):
        n = min(step, nbytes - o)
        hi = torch.randint(-7, 8, (n,), dtype=torch.int8, device=device, generator=gen)
        lo = torch.randint(-7, 8, (n,), dtype=torch.int8, device=device, generator=gen)
        out[o:o + n] = (hi << 4) | (lo & 0xF)
        del hi, lo
    return out


def host_random_nibbles(rng, nbytes):
    hi = rng.integers(-7, 8, nbytes, dtype=np.int8)
    lo = rng.integers(-7, 8, nbytes, dtype=np.int8)
    return ((hi.astype(np.uint8) << 4) | (lo.astype(np.uint8) & 0xF)).view(np.int8)


# ----------------------------------------------------------------------------------------------------------------
# the reference arm / cpu baseline: Clover's own mvm_parallel on the host cores
# ----------------------------------------------------------------------------------------------------------------
def host_threads():
    """all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1 to its workers - ignored on purpose)"""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def reference_cpu_gemv(sample_rows, cols, steps, warmup, full_rows=ROWS):
    """CloverMatrix4::mvm_parallel of the compiled reference (oracle/_ref) on all host threads; `sample_rows` rows of
    the `full_rows` x cols matrix (the default is the WHOLE matrix: same config as the GPU arm)."""
    threads = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(threads)          # before libgomp is loaded by the checker libraries
    os.environ.setdefault("OMP_PROC_BIND", "close")
    from oracle.pyoracle import Reference, aligned
    if Reference.available(False):
        ref, kind = Reference(False, threads=threads), "reference"
    else:
        ref, kind = None, "port"
    rng = np.random.default_rng(1234)
    # a 128 MiB block of uniform nibbles in [-7, 7], repeated: the reference's time does not depend on the values
    blk = min(sample_rows * cols // 2, 1 << 27)
    mv_blk = host_random_nibbles(rng, blk)
    ms = rng.uniform(0.25, 1.0, (sample_rows // 64) * (cols // 64)).astype(np.float32)
    xv = aligned(cols // 2, np.int8); xv[:] = host_random_nibbles(rng, cols // 2)
    xs = aligned(cols // 64, np.float32); xs[:] = rng.uniform(0.25, 1.0, cols // 64).astype(np.float32)
    nbytes = gemv_bytes(sample_rows, cols)
    times = []
    if ref is not None:
        m = Reference._M(ref, 4, sample_rows, cols)
        vals = m.values
        for o in range(0, vals.size, blk):
            n = min(blk, vals.size - o)
            vals[o:o + n] = mv_blk[:n]
        m.scales[:] = ms
        threads = ref.threads()
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            ref.m4_mvm(m, xv, xs, variant=2)             # CloverMatrix4::mvm_parallel (CloverMatrix4.h:1681)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        seq = []
        for i in range(2):                               # CloverMatrix4::mvm, one thread, for the record
            t0 = time.perf_counter()
            ref.m4_mvm(m, xv, xs, variant=0)
            seq.append(time.perf_counter() - t0)
        seq_note = f"; sequential mvm on 1 thread: {nbytes / min(seq) / 1e9:.2f} GB/s"
    else:
        from oracle.pyoracle import Oracle
        orc = Oracle()
        mv = np.tile(mv_blk, (sample_rows * cols // 2 + blk - 1) // blk)[: sample_rows * cols // 2]
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            orc.m4_mvm(mv, ms, sample_rows, cols, xv, xs)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    what = "the whole matrix" if sample_rows == full_rows else f"{sample_rows} of {full_rows} rows"
    return {"value": nbytes * len(times) / total / 1e9, "unit": "GB/s", "cores": threads, "kind": kind,
            "sample": f"{what} x {cols} cols, {len(times)} runs of CloverMatrix4::mvm_parallel on {threads} OpenMP threads "
                      f"(median {statistics.median(times) * 1e3:.2f} ms)" + (seq_note if ref is not None else ""),
            "ms_per_step": total / len(times) * 1e3, "same_config": sample_rows == full_rows}


def reference_cpu_extras():
    """The reference's own CPU routines for the other configs, timed beside the GPU numbers on bounded samples
    (quantize / dot / 8-bit mvm / scaleAndAdd / threshold, `_parallel` on all host threads and the
    sequential SIMD routine on one). GB/s by the reference's getBytes() model (01_measure.h:644,717,804,906)."""
    from oracle.pyoracle import Reference, aligned
    if not Reference.available(False):
        return {"unavailable": "oracle/_ref not built"}
    ref = Reference(False)
    rng = np.random.default_rng(7)
    out = {"cores": ref.threads(), "kind": "reference"}

    def best(fn, reps=3):
        ts = []
        for _ in range(reps + 1):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        return min(ts[1:])

    n = 1 << 24
    x = aligned(n, np.float32); x[:] = rng.uniform(-1, 1, n).astype(np.float32)
    v4 = lambda: (aligned(n // 2, np.int8), aligned(n // 64, np.float32))
    qv, qs = ref.v4_quantize(x, n)
    qv2, qs2 = ref.v4_quantize(x[::-1].copy(), n)
    vec_bytes = n // 2 + n // 64 * 4
    for name, var in (("parallel", 2), ("sequential", 0)):
        out[f"quantize4_n2^24_{name}_GBps"] = (4 * n + vec_bytes) / best(lambda: ref.v4_quantize(x, n, variant=var)) / 1e9
        out[f"dot4_n2^24_{name}_GBps"] = 2 * vec_bytes / best(lambda: ref.v4_dot(qv, qs, qv2, qs2, n, variant=var)) / 1e9
        out[f"scaleAndAdd4_n2^24_{name}_GBps"] = 3 * vec_bytes / best(lambda: ref.scale_and_add(4, qv, qs, qv2, qs2, 0.5, n, variant=var)) / 1e9
    nt, kt = 1 << 20, 1 << 14
    out["threshold4_n2^20_k2^14_sequential_ms"] = best(lambda: ref.threshold(4, qv[: nt // 2], qs[: nt // 64], nt, kt, variant=0)) * 1e3
    out["threshold4_n2^20_k2^14_parallel_ms"] = best(lambda: ref.threshold(4, qv[: nt // 2], qs[: nt // 64], nt, kt, variant=2)) * 1e3
    rows8, cols8 = 4096, 32768
    mv = rng.integers(-127, 128, rows8 * cols8, dtype=np.int8)
    ms = rng.uniform(0.25, 1.0, (rows8 // 64) * (cols8 // 64)).astype(np.float32)
    m8 = ref.m8_from(mv, ms, rows8, cols8)
    xv = aligned(cols8, np.int8); xv[:] = rng.integers(-127, 128, cols8, dtype=np.int8)
    xs = aligned(cols8 // 64, np.float32); xs[:] = rng.uniform(0.25, 1.0, cols8 // 64).astype(np.float32)
    b8 = gemv_bytes(rows8, cols8, 8)
    out["C5_mvm8_4096of32768rows_parallel_GBps"] = b8 / best(lambda: ref.m8_mvm(m8, xv, xs, variant=2)) / 1e9
    out["C5_mvm8_4096of32768rows_sequential_GBps"] = b8 / best(lambda: ref.m8_mvm(m8, xv, xs, variant=0), 2) / 1e9
    return {k: (round(v, 3) if isinstance(v, float) else v) for k, v in out.items()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = {"impl": "reference", "metric": METRIC, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int4 x int4 -> int32, fp32 scale epilogue", "data": "synthetic"}
    sample_rows = args.cpu_sample_rows or args.rows
    r = reference_cpu_gemv(sample_rows, args.cols, args.steps, args.warmup, args.rows)
    base.update({"value": r["value"], "ms_per_step": r["ms_per_step"],
                 "config": {"workload": f"CloverMatrix4::mvm {args.rows}x{args.cols} x CloverVector4 -> CloverVector4 (BASELINE C3)",
                            "rows": args.rows, "cols": args.cols, "sample_rows": sample_rows,
                            "reference_routine": "CloverMatrix4::mvm_parallel (include/CloverMatrix4.h:1681), unmodified reference compiled into oracle/_ref",
                            "same_config": r["same_config"]},
                 "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                 "e2e": {"value": r["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0})
    print(json.dumps(base), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# extras: the other BASELINE.json configs, CUDA-event timed
# ----------------------------------------------------------------------------------------------------------------
def cuda_time(torch, fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def extras(torch, cb, peak):
    from clover_b200 import DOT_EXACT, DOT_FAST
    out = {}
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev).manual_seed(99)
    # C2: quantize + dot at n = 2^26. dot operands (75.5 MB) fit in L2, so rotate over 4 operand sets.
    n = 1 << 26
    xs32 = []
    gen_key = np.zeros(8, np.uint64)
    import clover_b200 as _cbm
    _cbm.call("clover_prng_init", C.c_uint64(445560390295639063), C.c_uint64(2935984234003016713), gen_key.ctypes.data_as(C.c_void_p))
    for _ in range(2):                                   # the reference's generator and seeds
        v = cb.CloverVector32(n); v.setRandomFloats(-1.0, 1.0, gen_key); xs32.append(v)
    t = cuda_time(torch, lambda: xs32[0].setRandomFloats(-1.0, 1.0, gen_key), 10)
    out["setRandomFloats_n2^26"] = {"ms": t * 1e3, "GBps": n * 4 / t / 1e9, "note": "CloverVector32::setRandomFloats on the device, bit-identical to the reference's stream"}
    qs = [cb.CloverVector4(n) for _ in range(8)]
    for i, q in enumerate(qs):
        q.quantize(xs32[i % 2])
    i = [0]
    def quant():
        qs[i[0] % 8].quantize(xs32[i[0] % 2]); i[0] += 1
    t = cuda_time(torch, quant, 20)
    b = n * 4 + qs[0].getBytes()
    out["C2a_quantize4_n2^26"] = {"ms": t * 1e3, "GBps": b / t / 1e9, "frac_hbm": b / t / 1e9 / peak, "bytes": b,
                                  "traffic": ncu_traffic("k_vquantize4:2^26")}
    res = torch.empty(1, dtype=torch.float32, device=dev)
    def dot():
        k = i[0] % 4; qs[2 * k].dot_device(qs[2 * k + 1], res, DOT_FAST); i[0] += 1
    t = cuda_time(torch, dot, 40)
    b = 2 * qs[0].getBytes()
    out["C2b_dot4_n2^26_fast_rotating4"] = {"ms": t * 1e3, "GBps": b / t / 1e9, "frac_hbm": b / t / 1e9 / peak, "bytes": b,
                                            "traffic": ncu_traffic("k_vdot_fast4:2^26")}
    # threshold at n = 2^26 (radix select, 7 launches); bytes = the reference's own model 2*getBytes (01_measure.h:906)
    from clover_b200 import THRESHOLD_FAST
    def thr():
        qs[i[0] % 8].threshold(n // 64, THRESHOLD_FAST); i[0] += 1
    t = cuda_time(torch, thr, 8)
    out["threshold4_n2^26_k=2^20_fast"] = {"ms": t * 1e3, "GBps_ref_model": 2 * qs[0].getBytes() / t / 1e9}
    # scaleAndAdd (quantized AXPY) at n = 2^26; bytes = 3 * getBytes (01_measure.h:850)
    def axpy():
        k = i[0] % 2; qs[4 * k].scaleAndAdd(qs[4 * k + 1], 0.5, qs[4 * k + 2]); i[0] += 1
    t = cuda_time(torch, axpy, 20)
    b = 3 * qs[0].getBytes()
    out["scaleAndAdd4_n2^26"] = {"ms": t * 1e3, "GBps": b / t / 1e9, "frac_hbm": b / t / 1e9 / peak, "bytes": b,
                                 "traffic": ncu_traffic("scaleAndAdd4:2^26")}
    q8 = [cb.CloverVector8(n) for _ in range(2)]
    def quant8():
        q8[i[0] % 2].quantize(xs32[i[0] % 2]); i[0] += 1
    t = cuda_time(torch, quant8, 20)
    b = n * 4 + q8[0].getBytes()
    out["quantize8_n2^26"] = {"ms": t * 1e3, "GBps": b / t / 1e9, "frac_hbm": b / t / 1e9 / peak, "bytes": b}
    del xs32, qs, q8
    # C1: dot n = 4096 exact order (latency)
    a, b_ = cb.CloverVector4(4096), cb.CloverVector4(4096)
    v = cb.CloverVector32(4096); v.values.uniform_(-1, 1, generator=g); a.quantize(v)
    v.values.uniform_(-1, 1, generator=g); b_.quantize(v)
    t = cuda_time(torch, lambda: a.dot_device(b_, res, DOT_EXACT), 200)
    out["C1_dot4_n4096_exact"] = {"us": t * 1e6}
    # C5: 8-bit GEMV 32768^2
    r8 = c8 = 32768
    m8 = cb.CloverMatrix8(r8, c8)
    m8.values.copy_(torch.randint(-127, 128, (r8 * c8,), dtype=torch.int8, device=dev, generator=g))
    m8.scales.uniform_(0.25, 1.0, generator=g)
    x8, y8 = cb.CloverVector8(c8), cb.CloverVector8(r8)
    v = cb.CloverVector32(c8); v.values.uniform_(-1, 1, generator=g); x8.quantize(v)
    t = cuda_time(torch, lambda: m8.mvm(x8, y8), 20)
    b = gemv_bytes(r8, c8, 8)
    out["C5_gemv8_32768"] = {"ms": t * 1e3, "GBps": b / t / 1e9, "frac_hbm": b / t / 1e9 / peak, "bytes": b,
                             "traffic": ncu_traffic("k_m8_mvm_tma:32768x32768")}
    del m8
    # mixed precision, 4-bit matrix x CloverVector8 -> CloverVector8 at 32768^2 (CloverMatrix4.h:1093-1441)
    m4 = cb.CloverMatrix4(r8, c8)
    m4.values.copy_(random_nibbles(torch, r8 * c8 // 2, g, dev)); m4.scales.uniform_(0.25, 1.0, generator=g)
    t = cuda_time(torch, lambda: m4.mvm(x8, y8), 20)
    b = m4.getBytes() + x8.getBytes() + y8.getBytes()
    out["mvm4_v8_mixed_32768"] = {"ms": t * 1e3, "GBps": b / t / 1e9, "frac_hbm": b / t / 1e9 / peak, "bytes": b,
                                  "traffic": ncu_traffic("mvm4_v8_mixed:32768x32768")}
    # CloverMatrix4::mvm(V32,V32) - fp32 vectors, 32 chains per row (CloverMatrix4.h:1451-1547); bytes = matrix + x + y
    x32v, y32v = cb.CloverVector32(c8), cb.CloverVector32(r8)
    x32v.values.uniform_(-1, 1, generator=g)
    t = cuda_time(torch, lambda: m4.mvm(x32v, y32v), 20)
    b = m4.getBytes() + 4 * (c8 + r8)
    out["mvm4_f32_32768"] = {"ms": t * 1e3, "GBps": b / t / 1e9, "frac_hbm": b / t / 1e9 / peak, "bytes": b,
                             "fp32_ops_per_s": 2.0 * r8 * c8 / t,
                             "traffic": ncu_traffic("mvm4_f32:32768x32768"),
                             "note": "one int->float, one multiply and one fma per matrix element (the reference's order): CUDA-core bound, not HBM bound"}
    del m4
    # CloverMatrix8::mvm(V32,V32) (CloverMatrix8.h:558-661)
    m8 = cb.CloverMatrix8(r8, c8)
    m8.values.copy_(torch.randint(-127, 128, (r8 * c8,), dtype=torch.int8, device=dev, generator=g))
    m8.scales.uniform_(0.25, 1.0, generator=g)
    t = cuda_time(torch, lambda: m8.mvm(x32v, y32v), 20)
    b = m8.getBytes() + 4 * (c8 + r8)
    out["mvm8_f32_32768"] = {"ms": t * 1e3, "GBps": b / t / 1e9, "frac_hbm": b / t / 1e9 / peak, "bytes": b, "fp32_ops_per_s": 2.0 * r8 * c8 / t,
                             "traffic": ncu_traffic("mvm8_f32:32768x32768")}
    del m8
    # transpose of a 16384 x 16384 matrix (every byte read once and written once)
    for bits_, M_ in ((4, cb.CloverMatrix4), (8, cb.CloverMatrix8)):
        src, dst = M_(16384, 16384), M_(16384, 16384)
        src.values.copy_(torch.randint(-128, 128, (src.values.numel(),), dtype=torch.int8, device=dev, generator=g))
        src.scales.uniform_(0.25, 1.0, generator=g)
        t = cuda_time(torch, lambda: src.transpose(dst), 20)
        b = 2 * src.getBytes()
        out[f"transpose{bits_}_16384"] = {"ms": t * 1e3, "GBps": b / t / 1e9, "frac_hbm": b / t / 1e9 / peak, "bytes": b,
                                          "traffic": ncu_traffic(f"k_transpose{bits_}:16384x16384")}
        del src, dst
    # the reference's IHT loop (01_measure.h:924-946) at Phi 8192 x 32768, K = 1024: mvm, scaleAndAdd, mvm,
    # scaleAndAdd, threshold per iteration, 10 iterations per call, nothing read back
    from clover_b200 import apps
    Mi, Ni, Ki = 8192, 32768, 1024
    Phi, PhiT = cb.CloverMatrix4(Mi, Ni), cb.CloverMatrix4(Ni, Mi)
    Phi.values.copy_(torch.randint(-128, 128, (Phi.values.numel(),), dtype=torch.int8, device=dev, generator=g))
    Phi.scales.uniform_(0.01, 0.02, generator=g)
    Phi.transpose(PhiT)
    yv = cb.CloverVector4(Mi)
    v = cb.CloverVector32(Mi); v.values.uniform_(-1, 1, generator=g); yv.quantize(v)
    xv, t1, t2, t3 = cb.CloverVector4(Ni), cb.CloverVector4(Mi), cb.CloverVector4(Mi), cb.CloverVector4(Ni)
    t = cuda_time(torch, lambda: apps.Q_IHT(Phi, PhiT, xv, yv, t1, t2, t3, 10, Ki, 0.01, THRESHOLD_FAST), 5) / 10
    out["iht4_8192x32768_K1024_per_iteration"] = {"us": t * 1e6, "matrix_bytes": 2 * Phi.getBytes(), "GBps": 2 * Phi.getBytes() / t / 1e9,
                                                   "frac_hbm": 2 * Phi.getBytes() / t / 1e9 / peak}
    graph = apps.capture(lambda: apps.Q_IHT(Phi, PhiT, xv, yv, t1, t2, t3, 10, Ki, 0.01, THRESHOLD_FAST))
    t = cuda_time(torch, graph.replay, 5) / 10
    out["iht4_8192x32768_K1024_per_iteration_cuda_graph"] = {"us": t * 1e6, "GBps": 2 * Phi.getBytes() / t / 1e9,
                                                              "frac_hbm": 2 * Phi.getBytes() / t / 1e9 / peak}
    del Phi, PhiT
    # C4: 4-bit GEMM 16384^3
    M = N = K = 16384
    A, Bt = cb.CloverMatrix4(M, K), cb.CloverMatrix4(N, K)
    for m in (A, Bt):
        m.values.copy_(random_nibbles(torch, M * K // 2, g, dev)); m.scales.uniform_(0.25, 1.0, generator=g)
    Cout = torch.empty(M, N, dtype=torch.float32, device=dev)
    ops = 2.0 * M * N * K
    # whole call: nibble -> E4M3 expansion of both operands (HBM pass) + tcgen05 kernel; 20 back-to-back calls (sustained clocks)
    t = cuda_time(torch, lambda: A.gemm(Bt, out=Cout), 20, warmup=3)
    a8, b8 = A.expand_e4m3(), Bt.expand_e4m3()
    tk = cuda_time(torch, lambda: A.gemm_expanded(a8, Bt, b8, out=Cout), 20, warmup=3)
    out["C4_gemm4_16384"] = {
        "ms": t * 1e3, "TOPS": ops / t / 1e12, "kernel_only_ms": tk * 1e3, "kernel_only_TOPS": ops / tk / 1e12,
        "roofline": {"bound": "tensor", "kernel": "k_gemm4_tc", "achieved": ops / tk / 1e12, "peak": INT8_PEAK_TOPS,
                     "unit": "TOP/s", "frac": ops / tk / 1e12 / INT8_PEAK_TOPS,
                     "peak_source": "measured here: tools/mma_probe peak, tcgen05.mma kind::i8 issue loop on 148 SMs "
                                    "(profiles/r01_mma_probe.txt, re-measured profiles/r02_mma_probe_peak.txt: 4565 burst / 4572 sustained); "
                                    "kind::f8f6f4, the kind this kernel uses, sustains 3788",
                     "frac_of_e4m3_sustained_3767": ops / tk / 1e12 / 3767.0, "frac_of_nominal_4500": ops / tk / 1e12 / 4500.0,
                     "traffic": ncu_traffic("k_gemm4_tc:16384^3")},
        "note": "C[i][j] = rowView(A,i).dot(rowView(Bt,j)); bit-identical to the DP4A kernel (tests/test_gpu_parity.py)"}
    return out


# ----------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="clover_b200", choices=["clover_b200", "reference"])
    ap.add_argument("--rows", type=int, default=ROWS)
    ap.add_argument("--cols", type=int, default=COLS)
    ap.add_argument("--exchange", default="fused", choices=["fused", "fused_sync", "stamped", "allgather", "allreduce"])
    ap.add_argument("--cpu-sample-rows", type=int, default=0, help="rows of the matrix the CPU reference times (0 = all)")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "clover_b200" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import clover_b200
    from clover_b200 import containers as cb
    from clover_b200.sharded import ShardedCloverMatrix4

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = clover_b200.lib()
    peak, peak_src = measured_peaks()
    rows, cols = args.rows, args.cols

    # ---- operands: this rank's rows of the matrix, replicated x -------------------------------------------------
    # x: the reference's own generator and seeds (setRandomFloats(-1, 1), test/random/00_random.cpp:42), run on the
    # device by clover_v32_set_random_floats - the same x on every rank; the 2 GiB matrix is generated directly in its
    # quantized form (SURVEY.md 8d allows that for the perf-only runs, the quantizer is measured separately in C2)
    ref_key = np.zeros(8, np.uint64)
    clover_b200.call("clover_prng_init", C.c_uint64(445560390295639063), C.c_uint64(2935984234003016713), ref_key.ctypes.data_as(C.c_void_p))
    xf = cb.CloverVector32(cols)
    xf.setRandomFloats(-1.0, 1.0, ref_key)
    x = cb.CloverVector4(cols)
    x.quantize(xf)
    y = cb.CloverVector4(rows)
    A = ShardedCloverMatrix4(rows, cols, exchange=args.exchange)
    gm = torch.Generator(device=dev).manual_seed(1000 + rank)
    A.local.values[: A.rows_local * cols // 2].copy_(random_nibbles(torch, A.rows_local * cols // 2, gm, dev))
    A.local.scales.uniform_(0.25, 1.0, generator=gm)
    torch.cuda.synchronize()

    out = {"y": y}

    def step():
        if world == 1:
            A.local.mvm(x, y)          # the reference-facing call: CloverMatrix4::mvm(V4, V4), one fused kernel
        elif fused:
            out["y"] = A.mvm(x, wait=False)   # ONE kernel: shard GEMV + stamped NVLink-store epilogue (every result word reaches every
                                              # rank's message area); the unpack into the reference layout is the consumer's (finish())
        else:
            A.mvm(x, y)                # shard kernel + NCCL exchange + re-quantize

    fused = args.exchange in ("fused", "fused_sync", "stamped")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def finish():                      # fused exchange: the last step's result is complete on this rank (stream order)
        if world > 1 and fused:
            A.wait()

    total_bytes = gemv_bytes(rows, cols)
    for _ in range(args.warmup):
        step()
    finish()
    barrier()
    l0 = L.clover_kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        clk.mark()
        e0.record()
        for _ in range(args.steps):
            step()
        finish()
        e1.record()
        barrier()
        clk.mark()
    launches = L.clover_kernel_launches() - l0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    secs = float(ms.item()) * 1e-3
    value = total_bytes * args.steps / secs / 1e9

    # dominant kernel alone (the GEMV shard kernel), for the roofline: its own launches, CUDA events
    shard_bytes = gemv_bytes(A.rows_local, cols) if A.rows_local else 0
    def kernel_only():
        clover_b200.call("clover_m4_mvm_shard", C.c_void_p(A.local.values.data_ptr()), C.c_void_p(A.local.scales.data_ptr()),
                         C.c_uint64(A.rows_local), C.c_uint64(cols), C.c_uint64(A.row0), C.c_void_p(x.values.data_ptr()),
                         C.c_void_p(x.scales.data_ptr()), C.c_void_p(A.y32.data_ptr()),
                         C.c_void_p(y.values.data_ptr()), C.c_void_p(y.scales.data_ptr()), None,
                         C.c_void_p(torch.cuda.current_stream().cuda_stream))
    if world == 1:
        # one step IS one launch of this kernel and nothing else runs in the timed region: its average launch duration is the
        # timed region's own CUDA-event time (a second loop after 400 steps of full HBM load ran up to 5 % slower on warm boxes)
        tk = secs / args.steps
    else:
        tk = cuda_time(torch, kernel_only, args.steps)
    achieved = shard_bytes / tk / 1e9

    # ---- e2e: per-step operands in pinned host memory, result read back every step -------------------------------
    # Every byte moves through the product's own ABI: pinned buffers from clover_malloc_host;
    # N = 1: ONE call of clover_host_m4_mvm (x up, kernel, y down, sync - the host-buffer entry a reference user binds);
    # N > 1: clover_copy_h2d -> the sharded step -> clover_copy_d2h -> clover_stream_sync on the step's stream.
    # A container is one allocation [values | scales] (the reference's layout), so x goes up and y comes back in one copy each.
    call = clover_b200.call
    xb, yb = x.storage.numel(), y.storage.numel()
    hx, hy = C.c_void_p(), C.c_void_p()
    call("clover_malloc_host", C.byref(hx), C.c_size_t(xb))
    call("clover_malloc_host", C.byref(hy), C.c_size_t(yb))
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    call("clover_copy_d2h", hx, C.c_void_p(x.storage.data_ptr()), C.c_size_t(xb), stream)
    call("clover_stream_sync", stream)
    h2d, d2h = xb, yb
    xvb, yvb = x.values.numel(), y.values.numel()

    def e2e_step():
        if world == 1:
            call("clover_host_m4_mvm", C.c_void_p(A.local.values.data_ptr()), C.c_void_p(A.local.scales.data_ptr()),
                 C.c_uint64(rows), C.c_uint64(cols), hx, C.c_void_p(hx.value + xvb), hy, C.c_void_p(hy.value + yvb), None)
            return
        call("clover_copy_h2d", C.c_void_p(x.storage.data_ptr()), hx, C.c_size_t(xb), stream)
        if world > 1 and fused:
            out["y"] = A.mvm(x, wait=True)                # every step's result is read back: complete (unpacked) when the call ends
        else:
            step()
        r = out["y"]
        if getattr(r, "storage", None) is not None:
            call("clover_copy_d2h", hy, C.c_void_p(r.storage.data_ptr()), C.c_size_t(yb), stream)
        else:                                             # fused exchange: the result is a view into the shared block
            call("clover_copy_d2h", hy, C.c_void_p(r.values.data_ptr()), C.c_size_t(yvb), stream)
            call("clover_copy_d2h", C.c_void_p(hy.value + yvb), C.c_void_p(r.scales.data_ptr()), C.c_size_t(yb - yvb), stream)
        call("clover_stream_sync", stream)                # the caller reads the result of every step

    for _ in range(args.warmup):
        e2e_step()
    barrier()
    if world == 1:       # the host-buffer call returned the bytes the device-resident call produced
        assert C.string_at(hy.value, yb) == y.storage.cpu().numpy().tobytes(), "clover_host_m4_mvm result differs from clover_m4_mvm"
    e0.record()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = total_bytes * args.steps / (float(ms2.item()) * 1e-3) / 1e9

    # ---- north_star's wording, measured beside the fused exchange in the same run: the shard kernel,
    # ONE ncclAllReduce of the fp32 output, the re-quantize pass - same shards, same x, same timing protocol
    nccl = None
    unpack_each = None
    if world > 1 and fused:
        # the same loop with the result of EVERY step complete in the reference layout on every rank (mvm(x, wait=True))
        for _ in range(args.warmup):
            A.mvm(x, wait=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            out["y"] = A.mvm(x, wait=True)
        e1.record()
        barrier()
        ms4 = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(ms4, op=dist.ReduceOp.MAX)
        unpack_each = {"what": "every step's result complete in the reference layout on every rank (mvm(x, wait=True))",
                       "value": total_bytes * args.steps / (float(ms4.item()) * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": float(ms4.item()) / args.steps}
    if world > 1 and fused:
        B = ShardedCloverMatrix4.__new__(ShardedCloverMatrix4)
        B.__dict__.update(A.__dict__)
        B.exchange, B._peer = "allreduce", None
        y_ar = cb.CloverVector4(rows)
        for _ in range(args.warmup):
            B.mvm(x, y_ar)
        barrier()
        same = bool(torch.equal(y_ar.values, out["y"].values)) and bool(torch.equal(y_ar.scales.view(torch.int32)[: rows // 64], out["y"].scales.view(torch.int32)[: rows // 64]))
        e0.record()
        for _ in range(args.steps):
            B.mvm(x, y_ar)
        e1.record()
        barrier()
        ms3 = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(ms3, op=dist.ReduceOp.MAX)
        nccl = {"exchange": "shard kernel + one ncclAllReduce(fp32 y) + re-quantize", "value": total_bytes * args.steps / (float(ms3.item()) * 1e-3) / 1e9,
                "unit": "GB/s", "ms_per_step": float(ms3.item()) / args.steps, "same_bytes_as_fused": same}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": "int4 x int4 -> int32 (DP4A), fp32 scale epilogue", "data": "synthetic",
            "config": {"workload": f"CloverMatrix4::mvm {rows}x{cols} x CloverVector4 -> CloverVector4 (BASELINE C3)",
                       "rows": rows, "cols": cols, "algorithmic_bytes_per_step": total_bytes,
                       "rounding": "stochastic rounding disabled (parity configuration)",
                       "inputs": "x: the reference's setRandomFloats(-1, 1) stream from its fixed seeds (device generator), quantized; "
                                 "matrix: uniform nibbles in [-7, 7] and scales in [0.25, 1) generated on the device",
                       "l2_policy": "inputs larger than L2 (2 GiB matrix streamed per step vs 126 MB L2)",
                       "parallelism": "1 GPU" if world == 1 else
                                      (f"rows sharded over {world} GPUs in 64-row blocks; fused exchange: the GEMV epilogue stores every word of each "
                                       f"re-quantized block into every peer's message area over NVLink as one 8-byte {{word, epoch}} store (no NCCL call, no "
                                       f"fence, no flags); the unpack into the reference layout (clover_m4_shard_stamped_unpack) belongs to the consumer: "
                                       f"once, for the last step, inside the timed region (exchange.unpack_every_step: the same loop with it in every step)"
                                       if args.exchange in ("fused", "stamped") else
                                       f"rows sharded over {world} GPUs in 64-row blocks; fused exchange, flag form: plain peer stores, a system-scope fence "
                                       f"per CTA, one flag per peer, every kernel waits for the peers' flags at its end" if args.exchange == "fused_sync" else
                                       f"rows sharded over {world} GPUs in 64-row blocks + one NCCL {args.exchange} of the fp32 output"),
                       "e2e": ("clover_host_m4_mvm: x copied from pinned host memory, kernel, y copied back, host sync - every step; "
                               "matrix resident in HBM") if world == 1 else
                              "clover_copy_h2d(x) + sharded step (complete when its kernel ends) + clover_copy_d2h(y) + clover_stream_sync every step; "
                              "matrix shards resident in HBM"},
            "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": float(ms2.item()) / args.steps, "wall_ms_per_step": wall / args.steps * 1e3},
            "gpu_launches": launches,
            "clocks": clk.summary(),
            "exchange": {"mode": args.exchange if world > 1 else "none (1 GPU)",
                         "step_minus_kernel_us": (secs / args.steps - tk) * 1e6,      # what the exchange + launch gaps cost per step
                         "unpack_every_step": unpack_each,
                         "nccl_allreduce": nccl},
            "roofline": {"bound": "hbm", "kernel": "k_m4_mvm_tma2 (32-row items, 2 CTAs/SM)" if cols >= 16384 else "k_m4_mvm_tma", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_src,
                         "traffic": ncu_traffic(f"C3_mvm4:{rows}x{cols}") if world == 1 else None,
                         "kernel_ms": tk * 1e3, "algorithmic_bytes_per_launch": shard_bytes,
                         "note": "peak is the measured COPY bandwidth (a kernel that reads and writes); this kernel only reads, "
                                 "so frac may exceed 1 (ncu: 6.86 TB/s of DRAM traffic, profiles/r02u_gemv4_ncu_summary.txt)"},
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = {k: v for k, v in reference_cpu_gemv(args.cpu_sample_rows or rows, cols, 7, 2, rows).items()
                                        if k != "ms_per_step"}
            except Exception as exc:  # the checker failing must not hide the GPU number
                line["cpu_baseline"] = {"error": repr(exc)}
        if world == 1 and not args.no_extras:
            del A
            torch.cuda.empty_cache()
            try:
                line["extras"] = extras(torch, cb, peak)
            except Exception as exc:
                line["extras"] = {"error": repr(exc)}
            if not args.no_cpu_baseline:
                try:
                    line["extras"]["cpu_reference"] = reference_cpu_extras()
                except Exception as exc:
                    line["extras"]["cpu_reference"] = {"error": repr(exc)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
