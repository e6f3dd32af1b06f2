"""Summarise an ncu report (run here, no GPU needed): key metrics of each captured kernel.

usage: python tools/summarize_ncu.py gpurun_out/prof_X.ncu-rep [more.ncu-rep ...]
"""
import csv, io, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max",
    "smsp__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "sm__sass_inst_executed_op_shared_ld.sum",
    "smsp__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second",
]

def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(rep, "no data"); continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            print(f"== {rep} :: {d.get('Kernel Name','?')[:90]}")
            for k in KEYS:
                if k in d:
                    print(f"   {k:75s} {d[k]:>18s} {units[hdr.index(k)]}")
            def in_bytes(key):
                if key not in hdr:
                    return 0.0
                mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[units[hdr.index(key)]]
                return float(d[key].replace(",", "") or 0) * mult
            def in_seconds(key):
                mult = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "second": 1.0}[units[hdr.index(key)]]
                return float(d[key].replace(",", "") or 0) * mult
            traffic = in_bytes("dram__bytes_read.sum") + in_bytes("dram__bytes_write.sum")
            t = in_seconds("gpu__time_duration.sum")
            print(f"   DRAM traffic (read+write) = {traffic:.0f} B per launch; duration {t * 1e6:.2f} us under ncu "
                  f"-> {traffic / t / 1e9:.0f} GB/s of DRAM traffic")

if __name__ == "__main__":
    main()
