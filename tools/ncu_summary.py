"""Summarise an .ncu-rep (one kernel, --set full) as markdown: the headline metrics, the stall mix and the hottest
SASS lines.    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.md"""
import csv
import io
import subprocess
import sys


def page(rep, which):
    out = subprocess.run(["ncu", "-i", rep, "--page", which, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep):
    raw = page(rep, "raw")
    hdr, unit, val = raw[0], raw[1], raw[2]
    d = {h: (v, u) for h, u, v in zip(hdr, unit, val)}
    want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max",
            "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
            "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    print("| metric | value |\n|---|---|")
    for w in want:
        if w in d:
            print("| `%s` | %s %s |" % (w, d[w][0], d[w][1]))
    print("\nStall mix (`smsp__average_warps_issue_stalled_*_per_issue_active.ratio`, cycles per issued instruction):\n")
    print("| reason | ratio |\n|---|---|")
    stalls = []
    for h in hdr:
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                stalls.append((float(d[h][0].replace(",", "")), h.split("issue_stalled_")[1].split("_per_issue")[0]))
            except ValueError:
                pass
    for v, n in sorted(stalls, reverse=True)[:8]:
        print("| %s | %.2f |" % (n, v))
    src = page(rep, "source")
    if len(src) > 2:
        h2 = src[1]
        iS, iN, iI = h2.index("Source"), h2.index("# Samples"), h2.index("Instructions Executed")
        rows = [(int(r[iN] or 0), int(r[iI] or 0), r[iS].strip()) for r in src[2:] if len(r) > iI]
        tot = sum(r[0] for r in rows) or 1
        print("\nHottest SASS lines (warp-stall samples, share, executions):\n")
        print("| samples | share | executed | SASS |\n|---|---|---|---|")
        for n, i, s in sorted(rows, reverse=True)[:14]:
            print("| %d | %.1f %% | %d | `%s` |" % (n, 100.0 * n / tot, i, s[:90]))


if __name__ == "__main__":
    main(sys.argv[1])
