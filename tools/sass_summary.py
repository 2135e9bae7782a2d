#!/usr/bin/env python
"""Per-kernel SASS opcode summary of libmrag.so (no GPU needed: cuobjdump reads the cubin).

    python tools/sass_summary.py [profiles/r2_sass_opcodes.txt]

Counts, per kernel, the instructions that prove which hardware path the code is on
(B200_PROFILING.md): UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (TMEM), UTMALDG/UTMASTG = TMA
tensor loads/stores, UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier ops, HMMA = legacy mma.sync,
HGMMA = wgmma (must be 0 on sm_100a), LDG.E.128 = 128-bit global loads, plus registers per thread.
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "motionrag_b200" / "_lib" / "libmrag.so"
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "HGMMA", "LDG.E.128",
       "LDGSTS", "ATOM", "RED", "SHFL", "FFMA", "BAR.SYNC"]


def main(out=None):
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", str(LIB)], capture_output=True, text=True).stdout
    regs = {}
    fn = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
        m = re.search(r"REG:(\d+).*SHARED:(\d+)", line)
        if m and fn:
            regs[fn] = (int(m.group(1)), int(m.group(2)))
    counts = collections.OrderedDict()
    fn = None
    for line in sass.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
            counts[fn] = collections.Counter()
            continue
        if fn is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m:
            continue
        op = m.group(1)
        counts[fn]["_total"] += 1
        for key in OPS:
            if (op == key or op.startswith(key + ".") or (key == "LDG.E.128" and op.startswith("LDG.E") and ".128" in op)
                    or (key in ("ATOM", "RED") and op.startswith(key))):
                counts[fn][key] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    lines = [f"# SASS opcode summary of {LIB.relative_to(ROOT)} ({LIB.stat().st_size} bytes, {len(counts)} kernels); "
             "cuobjdump -sass, counted by tools/sass_summary.py",
             "# " + " ".join(f"{o:>9s}" for o in ["regs", "smem", "instrs"] + OPS) + "  kernel"]
    tot = collections.Counter()
    for (fn, c), name in zip(counts.items(), demangle):
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("mrag::", "")
        r, s = regs.get(fn, (0, 0))
        lines.append("  " + " ".join(f"{v:9d}" for v in [r, s, c["_total"]] + [c[o] for o in OPS]) + "  " + name)
        tot.update(c)
    lines.append("# totals: " + ", ".join(f"{o} {tot[o]}" for o in OPS))
    text = "\n".join(lines) + "\n"
    if out:
        Path(out).write_text(text)
    print(text)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
