#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: hottest SASS instructions with their dominant stall reasons."""
import csv, sys
path = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
for n, hi in enumerate(his):
    hdr = rows[hi]
    idx = {h: i for i, h in enumerate(hdr)}
    end = his[n + 1] - 1 if n + 1 < len(his) else len(rows)
    data = [r for r in rows[hi + 1:end] if len(r) > 10 and r[idx['# Samples']].isdigit()]
    name = rows[hi - 1][1] if hi > 0 else '?'
    tot = sum(int(r[idx['# Samples']]) for r in data) or 1
    print(f"=== {name}: {len(data)} SASS rows, {tot} samples")
    stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    agg = {h: sum(int(r[idx[h]] or 0) for r in data) for h in stall_cols}
    print('   stall totals:', sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
    for r in sorted(data, key=lambda r: -int(r[idx['# Samples']]))[:topn]:
        st = {h: int(r[idx[h]] or 0) for h in stall_cols}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print(f"{int(r[idx['# Samples']]):7d} {100*int(r[idx['# Samples']])/tot:5.1f}% exec={r[idx['Instructions Executed']]:>9} {r[idx['Source']][:64]:64s} {top}")
