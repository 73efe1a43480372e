"""Per-kernel SASS evidence: for every kernel of the built extension, the count of the Blackwell
mnemonics that prove which hardware path it uses (tcgen05 = UTCHMMA / LDTM / UTCBAR /
UTCATOMSWS, TMA = UTMALDG / UTMASTG / UTMACCTL, mbarrier = SYNCS, multimem/NVLS = LDGMC / REDG…MC
/ STG…MC, clusters = UCGABAR, PDL = ACQBULK / PREEXIT, cp.async = LDGSTS, vector atomics =
ATOMG…F32x4), plus the lines around the first tensor-core instruction.

    python tools/sass_evidence.py            # writes profiles/sass/summary.md + excerpts
"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = re.compile(r"\b(UTCHMMA|LDTM|UTCBAR|UTCATOMSWS|UTMALDG|UTMASTG|UTMACCTL|UTMAPF|SYNCS|LDGMC|"
                 r"REDGMC|STGMC|[A-Z]*MC\b|UCGABAR_ARV|UCGABAR_WAIT|ACQBULK|PREEXIT|LDGSTS|"
                 r"ATOMG|REDG|REDUX|MEMBAR|ERRBAR|HMMA|IMMA)[A-Za-z0-9_.]*")


def main() -> int:
    so = [f for f in os.listdir(os.path.join(ROOT, "ddlbench_b200", "ops")) if f.startswith("_ddlb_ext") and f.endswith(".so")]
    if not so:
        print("build the extension first (python -c 'import __graft_entry__ as g; g.build()')")
        return 1
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "ddlbench_b200", "ops", so[0])],
                          capture_output=True, text=True).stdout
    demangle = {}
    kernels: "OrderedDict[str, list]" = OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = []
        elif cur is not None:
            kernels[cur].append(line)
    names = list(kernels)
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    for n, d in zip(names, out):
        demangle[n] = re.sub(r"\(.*", "", d).replace("void ", "")
    os.makedirs(os.path.join(ROOT, "profiles", "sass"), exist_ok=True)
    rows = []
    for n, lines in kernels.items():
        c = Counter()
        first_mma = None
        for i, ln in enumerate(lines):
            for m in PAT.finditer(ln):
                c[m.group(0).rstrip(".")] += 1
            if first_mma is None and ("UTCHMMA" in ln or "LDGMC" in ln or "UCGABAR_ARV" in ln):
                first_mma = i
        rows.append((demangle[n], len([l for l in lines if re.search(r"/\*[0-9a-f]{4}\*/", l)]), c))
        if first_mma is not None:
            tag = re.sub(r"[^A-Za-z0-9_]+", "_", demangle[n])[:80]
            with open(os.path.join(ROOT, "profiles", "sass", tag + ".sass"), "w") as f:
                f.write("// %s\n// excerpt around the first tcgen05 / multimem / cluster instruction\n" % demangle[n])
                f.write("\n".join(l.rstrip() for l in lines[max(0, first_mma - 12):first_mma + 28]) + "\n")
    with open(os.path.join(ROOT, "profiles", "sass", "summary.md"), "w") as f:
        f.write("# SASS evidence (cuobjdump -sass of ddlbench_b200/ops/_ddlb_ext*.so, sm_100a)\n\n")
        f.write("| kernel | SASS lines | notable instructions (count) |\n|---|---|---|\n")
        for name, n, c in sorted(rows):
            keep = ", ".join("%s×%d" % (k, v) for k, v in sorted(c.items(), key=lambda kv: -kv[1])
                             if not k.startswith(("MEMBAR", "ERRBAR")))
            f.write("| `%s` | %d | %s |\n" % (name, n, keep or "–"))
    print("wrote profiles/sass/summary.md (%d kernels)" % len(rows))
    return 0


if __name__ == "__main__":
    sys.exit(main())
