"""Turn an .ncu-rep (ncu --set full) into a markdown summary and per-kernel DRAM traffic JSON.

usage: python tools/ncu_summarise.py <report.ncu-rep> <out.md> [<traffic.json> <npix_per_step> [<px_captured_launch> <launches_per_step> [note...]]]

If the captured launches cover <px_captured_launch> pixels each but the step's launches average
npix_per_step / launches_per_step pixels, the per-launch traffic is scaled by that ratio (traffic
is proportional to the pixels of a launch).
"""
import csv
import io
import json
import subprocess
import sys
from collections import defaultdict

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    rep, out_md = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    per_kernel = defaultdict(list)
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        per_kernel[name].append(r)
    traffic = {}
    with open(out_md, "w") as f:
        f.write(f"# ncu --set full --clock-control none summary of `{rep.split('/')[-1]}`\n\n")
        f.write(" ".join(sys.argv[7:]) + "\n\n" if len(sys.argv) > 7 else "")
        for name, rs in per_kernel.items():
            f.write(f"## {name} ({len(rs)} launches captured)\n\n| metric | " + " | ".join(f"launch {i}" for i in range(len(rs))) + " | unit |\n")
            f.write("|---|" + "---|" * (len(rs) + 1) + "\n")
            for w in WANT:
                if w in col:
                    f.write(f"| {w} | " + " | ".join(r[col[w]] for r in rs) + f" | {units[col[w]]} |\n")
            tot = 0.0
            for r in rs:
                for w in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    tot += float(r[col[w]].replace(",", "")) * UNIT.get(units[col[w]], 1.0)
            traffic[name] = {"launches": len(rs), "dram_bytes_total": tot, "dram_bytes_per_launch": tot / len(rs)}
            f.write(f"\nDRAM traffic (read+write) per launch, mean of the captured launches: {tot / len(rs) / 1e9:.3f} GB\n\n")
    if len(sys.argv) > 4:
        npix = int(sys.argv[4])
        scale = 1.0
        if len(sys.argv) > 6:
            scale = (npix / float(sys.argv[6])) / float(sys.argv[5])
        short = {"k_contract": "contract", "k_coeff<7>": "coeff", "k_coeff_oct<7>": "coeff", "k_irfft_czt": "irfft"}
        out = {}
        for name, t in traffic.items():
            for key, s in short.items():
                if name.startswith(key.split("<")[0]):
                    out[s] = {"npix": npix, "dram_bytes_per_launch": t["dram_bytes_per_launch"] * scale,
                              "dram_bytes_per_captured_launch": t["dram_bytes_per_launch"],
                              "launches_captured": t["launches"], "source": rep.split("/")[-1]}
        json.dump(out, open(sys.argv[3], "w"), indent=1)
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
