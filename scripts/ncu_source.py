#!/usr/bin/env python
"""Top CUDA source lines of a kernel by executed instructions and stall samples.
usage: ncu_source.py report.ncu-rep kernel-regex [top]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = [i for i, r in enumerate(rows) if 'Instructions Executed' in r][0]
hdr = rows[h]
ie, isamp = hdr.index('Instructions Executed'), hdr.index('# Samples')
src = [(int(r[ie]), int(r[isamp]) if r[isamp].isdigit() else 0, r[0], r[1]) for r in rows[h + 1:] if r[0] != '' and len(r) > ie and r[ie].isdigit()]
te, ts = sum(s[0] for s in src) or 1, sum(s[1] for s in src) or 1
print("total warp-instructions %d, samples %d" % (te, ts))
print("-- by stall samples")
for e, sm, ln, s in sorted(src, key=lambda x: -x[1])[:top]:
    print("%5.1f%% samp %5.1f%% inst  L%-4s %s" % (100 * sm / ts, 100 * e / te, ln, s.strip()[:100]))
print("-- by instructions")
for e, sm, ln, s in sorted(src, key=lambda x: -x[0])[:top]:
    print("%5.1f%% samp %5.1f%% inst  L%-4s %s" % (100 * sm / ts, 100 * e / te, ln, s.strip()[:100]))
