"""Summarise an .ncu-rep (read here, no GPU needed): python tools/ncu_summary.py <report.ncu-rep> [out.txt]

Prints, per profiled launch, the metrics bench.py's roofline and DESIGN.md quote: duration, DRAM bytes,
registers, occupancy limiters, pipe utilisation, shared-memory bank conflicts, and the top stall reasons.
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "launch__registers_per_thread",
    "launch__block_size",
    "launch__grid_size",
    "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg",
    "sm__cycles_elapsed.max",
    "smsp__inst_executed.sum",
    "smsp__thread_inst_executed.sum",
    "sm__inst_executed_pipe_fp64.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum",
    "sm__inst_executed_pipe_alu.sum",
    "sm__inst_executed_pipe_fma.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed_op_shared_ld.sum",
    "smsp__inst_executed_op_shared_st.sum",
    "smsp__inst_executed_op_global_ld.sum",
    "smsp__inst_executed_op_global_st.sum",
    "smsp__inst_executed_op_local_ld.sum",
    "smsp__inst_executed_op_local_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]
STALL = "smsp__average_warps_issue_stalled_"  # ..._per_issue_active.ratio / warp latency breakdown


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    lines = []
    for data in rows[2:]:
        rec = dict(zip(hdr, data))
        u = dict(zip(hdr, units))
        lines.append(f"== {rec.get('Kernel Name')}  grid {rec.get('Grid Size')} block {rec.get('Block Size')}")
        for k in KEYS:
            for h in hdr:
                if h == k or h.endswith("." + k):
                    lines.append(f"  {k:75s} {rec[h]:>18s} {u[h]}")
                    break
        stalls = []
        for h in hdr:
            if STALL in h and h.endswith("_per_warp_active.pct"):
                try:
                    stalls.append((float(rec[h].replace(",", "")), h.split(STALL)[1].replace("_per_warp_active.pct", "")))
                except ValueError:
                    pass
        if not stalls:
            for h in hdr:
                if "warps_issue_stalled" in h and h.endswith(".ratio"):
                    try:
                        stalls.append((float(rec[h].replace(",", "")), h.split("stalled_")[1]))
                    except ValueError:
                        pass
        stalls.sort(reverse=True)
        lines.append("  top stalls: " + ", ".join(f"{n}={v:.2f}" for v, n in stalls[:8]))
        # FP64 work of the launch from the sass op counters (rates per elapsed SMSP cycle, summed over the SMSPs)
        try:
            f = lambda op: float(rec[f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed"].replace(",", ""))
            cyc = float(rec["smsp__cycles_elapsed.avg"].replace(",", ""))
            insts = (f("dfma") + f("dmul") + f("dadd")) * cyc
            flops = (2.0 * f("dfma") + f("dmul") + f("dadd")) * cyc
            ms = float(rec["gpu__time_duration.sum"].replace(",", ""))
            ms = ms * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u["gpu__time_duration.sum"], 1.0)
            lines.append(f"  fp64: {insts:.4e} thread instructions (DFMA + DMUL + DADD), {flops:.4e} flops "
                         f"(DFMA = 2) -> {flops / (ms * 1e-3) / 1e12:.2f} TFLOP/s under ncu")
        except (KeyError, ValueError):
            pass
    text = "\n".join(lines)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")


if __name__ == "__main__":
    main()
