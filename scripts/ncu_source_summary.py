#!/usr/bin/env python
"""Summarises an `ncu --page source --csv` export of one kernel: executed instructions,
shared-memory wavefronts, L1 tag requests and L2 sectors per opcode, per unit of work
(warp iterations).  Usage: ncu_source_summary.py source.csv units [annotated_out.txt]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2])
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
cols = ['L1 Tag Requests Global', 'L1 Wavefronts Shared', 'L1 Wavefronts Shared Ideal',
        'L2 Theoretical Sectors Global']
mem_ops = ('LDS', 'STS', 'LDG', 'STG', 'LDGSTS', 'REDG', 'UBLKCP', 'SHFL', 'LDL', 'STL')
agg = collections.defaultdict(collections.Counter)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
total = 0
samples = collections.Counter()
annot = []
for k, r in enumerate(rows[2:]):
    if len(r) < len(hdr):
        continue
    src = re.sub(r'^@!?U?P\w+\s+', '', r[idx['Source']].strip())
    op = src.split()[0].rstrip(';')
    parts = op.split('.')
    base = '.'.join(parts[:2]) if parts[0] in mem_ops else parts[0]
    n = int(r[idx['Instructions Executed']])
    total += n
    agg[base]['n'] += n
    agg[base]['samples'] += int(r[idx['# Samples']])
    for c in cols:
        try:
            agg[base][c] += int(r[idx[c]])
        except ValueError:
            pass
    st = sorted(((int(r[idx[h]]), h[6:]) for h in stalls), reverse=True)[:2]
    for h in stalls:
        samples[h[6:]] += int(r[idx[h]])
    annot.append(f"{k:5d} {n / units:6.2f} {int(r[idx['# Samples']]):5d} "
                 f"{st[0][1]:>10s}:{st[0][0]:<5d} {st[1][1]:>10s}:{st[1][0]:<5d} {r[idx['Source']].strip()}")
print(rows[0][1][:160])
print(f"instructions per unit: {total / units:.1f}")
tot_s = sum(samples.values())
print("stall samples: " + "  ".join(f"{k} {100 * v / tot_s:.1f}%" for k, v in samples.most_common(9)))
print('%-14s %8s %7s %8s %8s %8s %8s' % ('op', 'instr', 'samp%', 'tagreq', 'smemwf', 'ideal', 'l2sect'))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]['n']):
    if v['n'] / units < 1.5:
        continue
    print('%-14s %8.1f %7.1f %8.1f %8.1f %8.1f %8.1f' % (
        k, v['n'] / units, 100 * v['samples'] / max(1, tot_s), v[cols[0]] / units,
        v[cols[1]] / units, v[cols[2]] / units, v[cols[3]] / units))
if len(sys.argv) > 3:
    open(sys.argv[3], 'w').write('\n'.join(annot))
