"""Markdown summary of an ncu report (the metrics DESIGN.md / VERDICT read): python tools/ncu_summary.py <file.ncu-rep> [title]"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct"]
STALLS = "smsp__average_warps_issue_stalled_"


def main():
    path = sys.argv[1]
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    if len(sys.argv) > 2:
        print("# " + sys.argv[2] + "\n")
    for r in rows[2:]:
        print("## %s\n" % r[idx["Kernel Name"]])
        print("| metric | value | unit |\n|---|---|---|")
        print("| Block Size | %s |  |\n| Grid Size | %s |  |" % (r[idx["Block Size"]], r[idx["Grid Size"]]))
        for w in WANT:
            if w in idx:
                print("| %s | %s | %s |" % (w, r[idx[w]], units[idx[w]]))
        stalls = [(float(r[i].replace(",", "")), h) for h, i in idx.items()
                  if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")]
        for v, h in sorted(stalls, reverse=True)[:7]:
            print("| %s | %f | inst |" % (h, v))
        print()


if __name__ == "__main__":
    main()
