"""Hot SASS instructions of an ncu report with their source-line context: python tools/sass_hot.py rep.ncu-rep [min_share]"""
import csv, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
for i, r in enumerate(rows):
    if r and r[0] == 'Address': hdr = r; start = i + 1; break
iS = hdr.index('# Samples'); iE = hdr.index('Instructions Executed'); iSrc = hdr.index('Source')
def toint(x):
    try: return int(x)
    except Exception: return 0
body = [r for r in rows[start:] if len(r) > iE]
ts = sum(toint(r[iS]) for r in body); te = sum(toint(r[iE]) for r in body)
print('samples', ts, 'warp instructions', te)
stall_cols = [j for j, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
for k, r in enumerate(body):
    if toint(r[iS]) > ts * thr:
        st = sorted(((toint(r[j]), hdr[j][6:]) for j in stall_cols), reverse=True)[:2]
        ctx = ' | '.join(b[iSrc].strip()[:38] for b in body[max(0, k - 3):k])
        print(f"{k:6d} smp {100*toint(r[iS])/ts:5.1f}% exe {toint(r[iE]):9d} {r[iSrc].strip()[:60]:60s} {st}   <- {ctx}")
