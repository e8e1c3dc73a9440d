"""Summarise an .ncu-rep (read here, no GPU needed): per kernel launch the metrics the roofline argument uses.
    python scripts/ncu_summary.py gpurun_out/x.ncu-rep [--stalls]  > profiles/x.summary.txt
"""
import csv
import io
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "dur_us"), ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr_MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("l1tex__t_sector_hit_rate.pct", "l1hit%"), ("lts__t_sector_hit_rate.pct", "l2hit%"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("smsp__inst_executed.sum", "warp_inst")]
STALL = "smsp__average_warp_latency_issue_stalled_"
STALL2 = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print("# %s" % rep)
    print("%-58s " % "kernel" + " ".join("%10s" % n for _, n in WANT))
    for r in rows[2:]:
        name = r[col["Kernel Name"]][:58]
        vals = []
        for m, n in WANT:
            v = r[col[m]] if m in col else ""
            try:
                f = float(v.replace(",", ""))
                if n.endswith("_MB") and units[col[m]].lower().startswith("byte"):
                    f /= 1e6
                if n == "dur_us" and units[col[m]] in ("ns", "nsecond"):
                    f /= 1e3
                vals.append("%10.2f" % f if f < 1e6 else "%10.3g" % f)
            except ValueError:
                vals.append("%10s" % v[:10])
        print("%-58s " % name + " ".join(vals))
        if "--stalls" in sys.argv:
            st = []
            for h, i in col.items():
                if (h.startswith(STALL) or h.startswith(STALL2)) and h.endswith("_per_warp_active.pct") is False and "not_issued" not in h:
                    try:
                        st.append((float(r[i].replace(",", "")), h.replace(STALL, "").replace(STALL2, "")))
                    except ValueError:
                        pass
            st.sort(reverse=True)
            print("    stalls: " + ", ".join("%s=%.2f" % (n, v) for v, n in st[:7]))


if __name__ == "__main__":
    main()
