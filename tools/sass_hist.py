"""Per-opcode and per-execution-count histogram of an ncu report's SASS page (ncu -i REP --page source --csv --print-source sass)."""
import collections, csv, re, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; data = rows[2:]
iS, iN, iE = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
tot_s = sum(int(r[iN] or 0) for r in data); tot_e = sum(int(r[iE] or 0) for r in data)
print('samples', tot_s, 'inst', tot_e, 'sass lines', len(data))
op = collections.Counter(); ops = collections.Counter()
for r in data:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[iS]); o = m.group(2).split('.')[0] if m else '?'
    op[o] += int(r[iE] or 0); ops[o] += int(r[iN] or 0)
for o, c in op.most_common(18):
    print(f"{o:10s} inst {c:10d} {100*c/tot_e:5.1f}%  samples {100*ops[o]/tot_s:5.1f}%")
b = collections.OrderedDict()
for i, r in enumerate(data):
    e = int(r[iE] or 0); s = int(r[iN] or 0)
    a = b.setdefault(e, [0, 0, 0, i]); a[0] += 1; a[1] += e; a[2] += s
for k, a in sorted(b.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"exec/instr {k:8d}  n_sass {a[0]:5d}  inst {100*a[1]/tot_e:5.1f}%  samples {100*a[2]/tot_s:5.1f}% first_idx {a[3]}")
if len(sys.argv) > 2:
    lo, hi = int(sys.argv[2]), int(sys.argv[3])
    for r in data[lo:hi]:
        print(f"{int(r[iE] or 0):9d} {int(r[iN] or 0):5d}  {r[iS][:120]}")
