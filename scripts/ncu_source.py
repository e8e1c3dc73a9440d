"""Per-SASS-line view of one kernel of an .ncu-rep (read here, no GPU): samples, executed count, stall reasons.
    python scripts/ncu_source.py gpurun_out/x.ncu-rep [min_samples] [kernel_index]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
min_s = int(sys.argv[2]) if len(sys.argv) > 2 else 40
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# split per kernel: a "Kernel Name" row starts each block, the next row is the header
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
b = blocks[which]
hdr = b["rows"][0]
col = {h: i for i, h in enumerate(hdr)}
KEYS = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
tot_s = tot_i = 0
out = []
for n, r in enumerate(b["rows"][1:]):
    if len(r) < len(hdr):
        continue
    s = int(r[col["# Samples"]] or 0)
    i = int(r[col["Instructions Executed"]] or 0)
    tot_s += s; tot_i += i
    st = {k.replace("stall_", ""): int(r[col[k]]) for k in KEYS if r[col[k]] not in ("0", "")}
    out.append((n, r[col["Source"]].strip()[:70], s, i, st, r[col["L1 Wavefronts Shared"]] if "L1 Wavefronts Shared" in col else "-", r[col["L1 Wavefronts Shared Ideal"]] if "L1 Wavefronts Shared Ideal" in col else "-"))
print("#", b["name"], "samples", tot_s, "warp instructions", tot_i)
for o in out:
    if o[2] >= min_s or any(t in o[1] for t in ("SYNCS.PHASE", "UTMALDG", "UBLKCP", "NANOSLEEP")):
        print("%4d %-70s smp=%-5d exec=%-8d %s wf=%s/%s" % o)
