"""Per-source-line instruction and stall-sample shares of one kernel from an .ncu-rep (needs -lineinfo builds and
`--import-source on` captures):  python tools/ncu_lines.py <rep> <kernel regex> [top N] [launch index]"""
import collections, csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{kern}", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = collections.OrderedDict()
fname, hdr = None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Line No":
        hdr = r; iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); continue
    if hdr and r[0].isdigit() and r[2] == "-":          # a CUDA source line (SASS rows carry an address)
        inst, smp = float(r[iI] or 0), float(r[iS] or 0)
        if inst or smp:
            k = (fname, int(r[0]))
            a = agg.setdefault(k, [0.0, 0.0, r[1].strip()[:100]])
            a[0] += inst; a[1] += smp
# inlined functions are attributed to BOTH the callee line and the call site; totals use lines of leaf code only (approx.:
# report shares relative to the max over files to stay meaningful)
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f"{len(agg)} lines; (sums count inlined code at callee and call sites) inst {ti:.3g} samples {ts:.3g}")
byfile = collections.defaultdict(lambda: [0.0, 0.0])
for (f, l), v in agg.items():
    byfile[f][0] += v[0]; byfile[f][1] += v[1]
for f, v in byfile.items():
    print(f"  {f:16s} inst {v[0]:.3g}  samples {v[1]:.3g}")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f:14s} {l:4d}  inst {v[0]:10.3g}  samples {v[1]:7.0f}   {v[2]}")
