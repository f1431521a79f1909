"""Summarise an ncu report: headline metrics + warp-stall breakdown per kernel launch (reads `ncu -i ... --csv`)."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",          # FP64 + DMMA pipe shared by the 4 SMSPs
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main():
    rep = sys.argv[1]
    hdr, units, rows = raw(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows:
        print("=== %s  grid %s block %s" % (r[idx["Kernel Name"]], r[idx.get("launch__grid_size", 0)], r[idx.get("launch__block_size", 0)]))
        for k in KEYS:
            if k in idx:
                print("  %-72s %s %s" % (k, r[idx[k]], units[idx[k]]))
        st = []
        for h, i in idx.items():
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        tot = sum(v for v, _ in st) or 1.0
        print("  warp-stall cycles per issued instruction: " + ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:8]))


if __name__ == "__main__":
    main()
