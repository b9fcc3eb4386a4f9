"""Summarise the per-instruction stall samples of one kernel from `ncu --page source --csv` output."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr, data = rows[h], [r for r in rows[h + 1:] if len(r) >= len(rows[h]) and r[0] != 'Address']
ix = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
tot = {s: 0 for s in stalls}
samples = 0
opc = {}
for r in data:
    n = int(r[ix['# Samples']] or 0)
    samples += n
    for s in stalls:
        tot[s] += int(r[ix[s]] or 0)
    parts = r[ix['Source']].split()
    op = parts[1] if parts and parts[0].startswith('@') and len(parts) > 1 else (parts[0] if parts else '?')
    opc[op] = opc.get(op, 0) + n
print('total samples', samples)
for s, v in sorted(tot.items(), key=lambda x: -x[1])[:10]:
    print(f'  {s:28s} {v:8d} {100 * v / max(samples, 1):5.1f}%')
print(sorted(opc.items(), key=lambda x: -x[1])[:20])
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    dom = max(stalls, key=lambda s: int(r[ix[s]] or 0))
    print(f"{r[ix['# Samples']]:>7s} {dom[6:]:16s} {r[ix['Source']][:100]}")
