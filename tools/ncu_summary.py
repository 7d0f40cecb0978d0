"""Summarise an .ncu-rep (raw page) into a short text table: per captured launch the
duration, DRAM traffic, throughput percentages, instruction counts and the top warp
stall reasons. Usage: python tools/ncu_summary.py report.ncu-rep [out.txt]"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/smem pipe % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instr"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"],
                         stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        out.append("== %s  (id %s)" % (r[idx["Kernel Name"]], r[idx["ID"]]))
        for key, label in WANT:
            if key in idx:
                out.append("  %-28s %s %s" % (label, r[idx[key]], units[idx[key]]))
        stalls = []
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith(
                    "_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[idx[h]]), h[len("smsp__average_warps_issue_stalled_"):
                                                       -len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        out.append("  stalls (warps per issue): " +
                   ", ".join("%s %.2f" % (n, v) for v, n in stalls[:6]))
    text = "\n".join(out) + "\n"
    sys.stdout.write(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)


if __name__ == "__main__":
    main()
