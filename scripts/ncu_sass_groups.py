"""Per-instruction executed counts of a kernel in an .ncu-rep (source page), grouped into runs of equal count.
    python scripts/ncu_sass_groups.py gpurun_out/x.ncu-rep [dump.txt]"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]; ci = hdr.index('Instructions Executed'); si = hdr.index('Source'); ss = hdr.index('# Samples')
out = []
for r in rows[2:]:
    try: out.append((int(r[ci]), int(r[ss]), r[si].strip()))
    except (ValueError, IndexError): pass
if len(sys.argv) > 2:
    with open(sys.argv[2], 'w') as f:
        for i, (n, s, src) in enumerate(out): f.write("%d %d %d %s\n" % (i, n, s, src))
tot = sum(n for n, _, _ in out); samples = sum(s for _, s, _ in out)
print("total warp instructions %d, samples %d" % (tot, samples))
grp = []
for i, (n, s, src) in enumerate(out):
    if grp and grp[-1][2] == n: grp[-1][1] = i; grp[-1][3] += 1; grp[-1][4] += s
    else: grp.append([i, i, n, 1, s])
for g in grp:
    if g[2] * g[3] > 0.004 * tot or g[4] > 0.01 * samples:
        print("%4d-%4d count=%9d n=%3d inst=%9d (%4.1f%%) samples=%5d (%4.1f%%)  %s" % (g[0], g[1], g[2], g[3], g[2] * g[3], 100 * g[2] * g[3] / tot, g[4], 100 * g[4] / samples, out[g[0]][2][:40]))
