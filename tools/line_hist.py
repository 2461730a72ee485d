"""Per-source-line instruction / stall-sample share of an ncu report (needs -lineinfo and --import-source on)."""
import csv, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.005
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = {}; fname = ''
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; iE = r.index('Instructions Executed'); iN = r.index('# Samples'); continue
    if hdr is None or len(r) <= iE or not r[0].isdigit(): continue
    k = (fname, int(r[0]))
    a = agg.setdefault(k, [0, 0, r[1]])
    if not r[iE].isdigit(): continue
    a[0] += int(r[iE] or 0); a[1] += int(r[iN] or 0)
tot = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print('total inst', tot, 'samples', ts)
for k, a in sorted(agg.items()):
    if a[0] > tot * thr or a[1] > ts * thr:
        print(f"{k[0][:12]:12s}:{k[1]:4d} inst {100*a[0]/tot:5.1f}%  smp {100*a[1]/ts:5.1f}%  {a[2].strip()[:110]}")
