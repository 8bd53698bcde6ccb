#!/usr/bin/env python
"""Joins an `ncu --page source --csv` export of one kernel with `nvdisasm
--print-line-info-inline` of the same cubin: executed instructions and stall samples per
region of evaluate_kernel.cuh (outermost, i.e. not-inlined, source line of every SASS
instruction).  Usage:
  sass_by_source.py source.csv disasm.txt '<mangled kernel name>' units [regions]
regions: comma separated  name:first-last  over lines of the kernel's own file."""
import collections
import csv
import re
import sys

src_csv, disasm, kernel, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
regions = []
if len(sys.argv) > 5:
    for item in sys.argv[5].split(','):
        name, span = item.split(':')
        lo, hi = span.split('-')
        regions.append((name, int(lo), int(hi)))

BODY_FILE = 'evaluate_kernel.cuh'
BODY_LO = min([lo for _, lo, _ in regions] or [0])
BODY_HI = max([hi for _, _, hi in regions] or [10 ** 9])
# --- nvdisasm: per instruction (opcode text, outermost line, innermost file:line)
instrs = []
inside = False
pending = []
for line in open(disasm, errors='replace'):
    if line.startswith('\t.section\t.text.'):
        inside = line.startswith('\t.section\t.text.' + kernel + ',')
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)( inlined at)?', line)
    if m:
        pending.append((m.group(1), int(m.group(2)), bool(m.group(3))))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
    if m:
        if pending:
            inner = pending[0]
            # deepest frame that lies in the kernel body (helpers defined above the kernel
            # and library headers are attributed to the line that calls them)
            body_frames = [p for p in pending
                           if p[0].endswith(BODY_FILE) and BODY_LO <= p[1] <= BODY_HI]
            outer = body_frames[0] if body_frames else pending[-1]
            cur = (inner[0].split('/')[-1], inner[1], outer[0].split('/')[-1], outer[1])
            pending = []
        instrs.append((m.group(2).strip(), cur))

rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) >= len(hdr)]
if len(body) != len(instrs):
    print(f"warning: {len(body)} profiled instructions, {len(instrs)} disassembled", file=sys.stderr)

by_region = collections.defaultdict(lambda: [0, 0, collections.Counter()])
by_inner = collections.defaultdict(lambda: [0, 0])
total_n = total_s = 0
def stem(t):
    return re.sub(r'^@!?U?P\w+\s+', '', t.strip()).split()[0].rstrip(';').split('.')[0]
mismatch = sum(1 for r, (text, _) in zip(body, instrs) if stem(r[idx['Source']]) != stem(text))
if mismatch:
    print(f"warning: {mismatch} opcode mismatches between profile and disassembly", file=sys.stderr)
for r, (text, (ifile, iline, ofile, oline)) in zip(body, instrs):
    n = int(r[idx['Instructions Executed']])
    s = int(r[idx['# Samples']])
    total_n += n
    total_s += s
    name = f"{ofile}:{oline}"
    for rn, lo, hi in regions:
        if ofile.startswith('evaluate_kernel') and lo <= oline <= hi:
            name = rn
            break
    op = re.sub(r'^@!?U?P\w+\s+', '', text).split()[0].split('.')[0]
    by_region[name][0] += n
    by_region[name][1] += s
    by_region[name][2][op] += n
    by_inner[f"{ifile}:{iline}"][0] += n
    by_inner[f"{ifile}:{iline}"][1] += s
print(f"instructions per unit {total_n / units:.1f}, samples {total_s}")
print('%-28s %8s %7s  %s' % ('region', 'instr', 'samp%', 'top opcodes (instr per unit)'))
for name, (n, s, ops) in sorted(by_region.items(), key=lambda kv: -kv[1][1]):
    if n / units < 0.5 and s * 200 < total_s:
        continue
    top = ' '.join(f"{o}:{c / units:.0f}" for o, c in ops.most_common(7))
    print('%-28s %8.1f %7.1f  %s' % (name, n / units, 100.0 * s / max(1, total_s), top))
print()
print('innermost source lines by samples')
for name, (n, s) in sorted(by_inner.items(), key=lambda kv: -kv[1][1])[:40]:
    print('%-40s %8.1f %7.1f' % (name, n / units, 100.0 * s / max(1, total_s)))
