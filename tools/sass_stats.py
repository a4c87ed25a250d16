"""Per-kernel SASS statistics of a cubin/object/executable: registers, stack (spill) bytes, and
how many DFMAs take an operand from the operand-reuse cache.  Development aid (no GPU needed).

    python tools/sass_stats.py <binary> [name-filter]
"""
import re
import subprocess
import sys


def main():
    binary = sys.argv[1]
    filt = sys.argv[2] if len(sys.argv) > 2 else ""
    res = subprocess.run(["cuobjdump", "-res-usage", binary], capture_output=True, text=True).stdout
    usage = {}
    name = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        if name and "REG:" in line:
            usage[name] = dict(re.findall(r"(\w+):(\d+)", line))
            name = None
    sass = subprocess.run(["cuobjdump", "-sass", binary], capture_output=True, text=True).stdout
    stats = {}
    cur = None
    prev_reuse = set()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            stats[cur] = dict(dfma=0, cached=0, dadd=0, dmma=0, lds=0, ldg=0, total=0)
            prev_reuse = set()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", line)
        if not m:
            continue
        ins = m.group(1)
        st = stats[cur]
        st["total"] += 1
        op = ins.split()[1] if ins.startswith("@") else ins.split()[0]
        if op.startswith("DFMA") or op.startswith("DADD") or op.startswith("DMUL"):
            ops = [o.strip() for o in ins.split(None, 1)[1].split(",")] if not ins.startswith("@") else \
                  [o.strip() for o in ins.split(None, 2)[2].split(",")]
            srcs = ops[1:]
            # an operand is served by the reuse cache if the previous FP64 instruction flagged the
            # same register in the same slot
            hit = False
            flagged = set()
            for slot, o in enumerate(srcs):
                reg = re.sub(r"[-|]", "", o).replace(".reuse", "")
                if (slot, reg) in prev_reuse:
                    hit = True
                if ".reuse" in o:
                    flagged.add((slot, reg))
            prev_reuse = flagged
            if op.startswith("DFMA"):
                st["dfma"] += 1
                st["cached"] += hit
            else:
                st["dadd"] += 1
        else:
            if not (op.startswith("LDS") or op.startswith("LDG") or op.startswith("IMAD") or op.startswith("IADD")
                    or op.startswith("MOV") or op.startswith("LEA")):
                pass
            if op.startswith("DMMA"):
                st["dmma"] += 1
            if op.startswith("LDS"):
                st["lds"] += 1
            if op.startswith("LDG"):
                st["ldg"] += 1
    for k, st in stats.items():
        if filt and filt not in k:
            continue
        u = usage.get(k, {})
        print(f"{k[:70]:70s} REG {u.get('REG', '?'):>3} STACK {u.get('STACK', '?'):>4} "
              f"DFMA {st['dfma']:4d} (reuse-cached {st['cached']:4d}) DADD {st['dadd']:3d} DMMA {st['dmma']:4d} "
              f"LDS {st['lds']:3d} LDG {st['ldg']:3d} total {st['total']}")


if __name__ == "__main__":
    main()
