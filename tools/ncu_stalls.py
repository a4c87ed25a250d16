"""Stall breakdown of one kernel of an .ncu-rep: per-reason warp-stall ratios from the raw page and the
hottest instructions of the source page (ncu --set full --import-source on).
usage: python tools/ncu_stalls.py <report.ncu-rep> [top_n [kernel_regex]]   (first matching launch)"""
import csv
import io
import subprocess
import sys


KERNEL = []


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"] + KERNEL, capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    if len(sys.argv) > 3:
        KERNEL.extend(["-k", "regex:" + sys.argv[3], "-c", "1"])
    rows = page(rep, "raw")
    hdr, r = rows[0], rows[2]
    for h, v in zip(hdr, r):
        if h in ("gpu__time_duration.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
                 "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
                 "launch__registers_per_thread", "smsp__average_warps_active_per_inst_executed.ratio"):
            print(f"{h}: {v}")
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and float(v or 0) > 0.05:
            print(f"  {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]}: {float(v):.2f}")
    rows = page(rep, "source")
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for idx, r in enumerate(rows[2:]):
        try:
            n = float(r[ci["# Samples"]] or 0)
        except (ValueError, IndexError):
            continue
        why = max(stall_cols, key=lambda h: float(r[ci[h]] or 0))
        data.append((n, idx, why, r[ci["Source"]].strip()[:90]))
    tot = sum(d[0] for d in data)
    print(f"instructions {len(data)}, samples {int(tot)}")
    for n, idx, why, src in sorted(data, reverse=True)[:top_n]:
        print(f"  {idx:5d} {100 * n / tot:5.1f}%  {why:22s} {src}")


if __name__ == "__main__":
    main()
