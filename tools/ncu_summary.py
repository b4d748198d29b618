"""Summarise an `ncu --page raw --csv` export: one block of selected metrics per distinct kernel (first captured launch).
usage: python tools/ncu_summary.py gpurun_out/<tag>_full_raw.csv [> profiles/<tag>_ncu_summary.txt]"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__waves_per_multiprocessor",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "dram__sectors_read.sum", "dram__sectors_write.sum"]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    rows = list(csv.reader(open(path, newline="")))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    col = {n: i for i, n in enumerate(names)}
    seen = {}
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        k = r[col["Kernel Name"]].split("(")[0]
        seen.setdefault(k, []).append(r)
    for k, rs in seen.items():
        r = rs[0]
        print(f"===== {k}   ({len(rs)} captured launches; first shown)")
        for key in KEYS:
            if key in col and r[col[key]] != "":
                print(f"   {key:<75} {r[col[key]]} {units[col[key]]}")
        st = []
        for n, i in col.items():
            if n.startswith(STALL) and n.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a"):
                try:
                    st.append((float(r[i].replace(",", "")), n[len(STALL):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        st.sort(reverse=True)
        print("   stalls (warps stalled per issue-active cycle): " + ", ".join(f"{n}={v:.2f}" for v, n in st[:7]))
        if len(rs) > 1 and "gpu__time_duration.sum" in col:
            print("   durations of all captured launches: " + ", ".join(x[col["gpu__time_duration.sum"]] for x in rs) + " " + units[col["gpu__time_duration.sum"]])


if __name__ == "__main__":
    main(sys.argv[1])
