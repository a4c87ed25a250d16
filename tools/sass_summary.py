"""SASS opcode summary per kernel of libpkm_b200.so (no GPU needed): which Blackwell mechanisms each
hand-written kernel actually uses.  `python tools/sass_summary.py > profiles/r2_sass_summary.md`"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "popkinmocks_b200", "csrc", "libpkm_b200.so")
KEYS = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "DMMA", "DFMA", "DADD", "DMUL", "LDCU", "REDUX", "LDG", "STG",
        "LDS", "STS", "LDGSTS", "ATOM", "RED", "SHFL", "BAR", "UTCMMA", "LDTM", "HMMA"]


def demangle(names):
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    out = [o.replace("(int)", "").replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
           for o in out]
    return [re.sub(r"\(.*", "", o) for o in out]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs, name = {}, None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
        elif name and "REG:" in line:
            regs[name] = dict(re.findall(r"(\w+):(\d+)", line))
            name = None
    stats, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            stats[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", line)
        if cur is None or not m:
            continue
        ins = m.group(1).split()
        op = ins[1] if ins[0].startswith("@") else ins[0]
        stats[cur][op.split(".")[0]] += 1
        if op.startswith("DFMA") and "UR" in m.group(1):
            stats[cur]["DFMA_UR"] += 1
        stats[cur]["total"] += 1
    names = list(stats)
    pretty = demangle(names)
    print("# SASS opcode summary of libpkm_b200.so (sm_100a), per kernel\n")
    print("`cuobjdump -sass` of the shipped library, summarised by `tools/sass_summary.py`.  TMA = `UTMALDG` (tensor load), "
          "`UTMASTG` (tensor store), `UBLKCP` (1-D bulk copy); `SYNCS` = mbarrier operations; `DMMA` = FP64 tensor-core MMA "
          "(m8n8k4); `DFMA_UR` = DFMAs taking an operand from a uniform register (the contraction's spline weights, `LDCU` "
          "from the kernel-parameter bank).  tcgen05 (`UTC*MMA`, `LDTM`) has no FP64 kind, so none is expected.\n")
    cols = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "DMMA", "DFMA", "DFMA_UR", "LDCU", "DADD", "LDS", "LDG", "STG", "ATOM", "RED", "SHFL", "BAR"]
    print("| kernel | regs | instr | " + " | ".join(cols) + " |")
    print("|---|---|---|" + "---|" * len(cols))
    seen = set()
    for n, p in zip(names, pretty):
        if p in seen:
            continue
        seen.add(p)
        c = stats[n]
        print(f"| `{p}` | {regs.get(n, {}).get('REG', '?')} | {c['total']} | " + " | ".join(str(c.get(k, 0)) for k in cols) + " |")
    tot = collections.Counter()
    for c in stats.values():
        tot.update(c)
    print("\nWhole library: " + ", ".join(f"{k} x{tot[k]}" for k in ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "DMMA", "DFMA", "DFMA_UR", "LDCU", "UTCMMA", "LDTM", "HMMA"]) + ".")


if __name__ == "__main__":
    main()
