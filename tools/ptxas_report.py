#!/usr/bin/env python
"""Resource usage of every kernel in the extension from `cuobjdump --dump-resource-usage` on the
built objects (registers, spills via STACK, shared memory, constant bank) — the numbers
`-Xptxas -v` prints, without recompiling.  Writes profiles/resource_usage.md.
Usage: python tools/ptxas_report.py [out.md]"""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
BUILD = ROOT / "build" / "ddlb_ext"


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout
    return out.strip().split("\n")


def main(argv):
    out = Path(argv[0]) if argv else ROOT / "profiles" / "resource_usage.md"
    rows = []
    for obj in sorted(BUILD.glob("*.o")):
        txt = subprocess.run(["cuobjdump", "--dump-resource-usage", str(obj)], capture_output=True, text=True).stdout
        for m in re.finditer(r"Function (\S+):\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+) CONSTANT\[0\]:(\d+)", txt):
            rows.append((obj.stem.split(".")[0], m.group(1)) + tuple(int(g) for g in m.groups()[1:]))
    names = demangle([r[1] for r in rows])
    lines = ["# Kernel resource usage (sm_100a, from the built objects)", "",
             "`STACK` > 0 would mean register spills; dynamic shared memory (the tcgen05 kernels' rings) is",
             "requested at launch and not listed here.", "",
             "| file | kernel | regs | stack | static smem | local | const[0] |", "|---|---|---|---|---|---|---|"]
    for (f, _, reg, stack, sh, loc, c0), name in zip(rows, names):
        name = re.sub(r"\(.*", "", name.replace("void ", "")).replace("|", "\\|")
        lines.append("| %s | `%s` | %d | %d | %d | %d | %d |" % (f, name[:110], reg, stack, sh, loc, c0))
    out.write_text("\n".join(lines) + "\n")
    spills = [n for r, n in zip(rows, names) if r[3] > 0]
    print("%d kernels, %d with a stack frame -> %s" % (len(rows), len(spills), out))
    return 0


if __name__ == "__main__":
    raise SystemExit(main(sys.argv[1:]))
