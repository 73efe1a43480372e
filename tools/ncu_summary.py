"""Markdown summary of ncu reports: python tools/ncu_summary.py a.ncu-rep b.ncu-rep > out.md"""
import csv
import io
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time us"), ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem KB"),
        ("launch__occupancy_limit_registers", "CTA/SM (regs)"),
        ("launch__occupancy_limit_shared_mem", "CTA/SM (smem)"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("dram__bytes_read.sum", "DRAM read MB"), ("dram__bytes_write.sum", "DRAM write MB"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
        ("smsp__inst_executed.sum", "warp instructions"),
        ("launch__cluster_max_active", "max active clusters")]


def main(paths):
    for p in paths:
        raw = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr = rows[0]
        print("### %s\n" % p.split("/")[-1])
        stall = [(i, h) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
            print("**%s** grid %s block %s\n" % (name, r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
            print("| metric | value |\n|---|---|")
            for k, label in KEYS:
                if k in hdr and r[hdr.index(k)] not in ("", "0", "n/a"):
                    v = r[hdr.index(k)]
                    try:
                        v = "%.1f" % float(v.replace(",", ""))
                    except ValueError:
                        pass
                    print("| %s | %s |" % (label, v))
            top = sorted(((float(r[i] or 0), h.split("stalled_")[1].replace("_per_issue_active.ratio", ""))
                          for i, h in stall), reverse=True)[:5]
            if top:
                print("| top stalls (warps per issue) | %s |" % ", ".join("%s %.2f" % (n, v) for v, n in top))
            print()


if __name__ == "__main__":
    main(sys.argv[1:])
