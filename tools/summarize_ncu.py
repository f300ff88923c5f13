"""Turn .ncu-rep captures into the small text summaries committed under profiles/ (the reports themselves are scratch)."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max.per_second",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    lines = [f"# {rep} (ncu --set full --clock-control none; one launch)"]
    for i, h in enumerate(hdr):
        if h == "Kernel Name" or h in WANT:
            lines.append(f"{h}: {vals[i]} {units[i]}".rstrip())
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    if len(srows) > 2:
        h = srows[0]
        try:
            ci, si = h.index("Source"), h.index("# Samples") if "# Samples" in h else h.index("Sampling Data (All)")
            top = sorted((r for r in srows[1:] if len(r) > si and r[si].replace('.', '', 1).isdigit()), key=lambda r: -float(r[si]))[:12]
            lines.append("## hottest instructions (PC samples)")
            lines += [f"{r[si]:>8}  {r[ci][:140]}" for r in top]
        except ValueError:
            pass
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:24]))


if __name__ == "__main__":
    main()
