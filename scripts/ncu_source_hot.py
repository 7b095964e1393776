"""Hot SASS instructions of an `ncu --page source --csv --print-source sass` export: samples and top stall reasons.
usage: python scripts/ncu_source_hot.py src.csv [top N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 45
kern = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {'name': r[1], 'rows': []}; kern.append(cur); continue
    if cur is None: continue
    if r and r[0] == "Address": cur['hdr'] = r; continue
    if r: cur['rows'].append(r)
for k in kern:
    h = k['hdr']; ix = {n: i for i, n in enumerate(h)}
    R = k['rows']
    tot = sum(int(r[ix['# Samples']]) for r in R)
    print(k['name'][:40], 'instrs', len(R), 'samples', tot)
    order = sorted(range(len(R)), key=lambda i: -int(R[i][ix['# Samples']]))[:N]
    stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
    for i in sorted(order):
        r = R[i]; s = int(r[ix['# Samples']])
        top = sorted(((int(r[ix[n]]), n) for n in stalls), reverse=True)[:3]
        print(f"{i:6d} {r[ix['Source']].strip()[:60]:60s} {s:7d} {100 * s / tot:5.1f}%  exec {r[ix['Instructions Executed']]:>8s} ", ' '.join(f"{n[6:]}={v}" for v, n in top if v))
