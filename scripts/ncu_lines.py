"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line:
warp-instructions executed and stall samples. Usage: ncu_lines.py mix.csv [top] [a-b,c-d line ranges of the .cu]"""
import csv, sys, collections, os
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
hdr = None
cur, src, fname = None, {}, '?'
inst, samp = collections.Counter(), collections.Counter()
for r in rows:
    if r and r[0] == 'File Path':
        fname = os.path.basename(r[1]); continue
    if r and r[0] == 'Line No':
        hdr = r; iexe, isamp = hdr.index('Instructions Executed'), hdr.index('# Samples'); continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    if r[0]:
        cur = (fname, int(r[0])); src[cur] = r[1]
    elif cur is not None:
        try:
            inst[cur] += int(r[iexe]); samp[cur] += int(r[isamp])
        except ValueError:
            pass
ti, ts = sum(inst.values()), sum(samp.values())
print(f'total warp-instructions {ti}, samples {ts}')
byfile = collections.Counter()
for (f, l), n in inst.items(): byfile[f] += n
print({f: round(100 * n / ti, 1) for f, n in byfile.items()})
for key, n in sorted(inst.items(), key=lambda kv: -kv[1])[:top]:
    print(f'{key[0][:14]:14s}{key[1]:5d} {100*n/ti:5.1f}% inst {100*samp[key]/max(ts,1):5.1f}% samp  {src[key][:100]}')
if len(sys.argv) > 3:
    for rg in sys.argv[3].split(','):
        a, b = map(int, rg.split('-'))
        n = sum(v for k, v in inst.items() if k[0].endswith('.cu') and a <= k[1] <= b)
        s = sum(v for k, v in samp.items() if k[0].endswith('.cu') and a <= k[1] <= b)
        print(f'lines {a}-{b}: {100*n/ti:5.1f}% inst, {100*s/max(ts,1):5.1f}% samples')
