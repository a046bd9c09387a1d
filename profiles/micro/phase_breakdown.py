"""Folds an `ncu --set full --import-source on` capture of one kernel per barrier-separated phase.

    ncu -i report.ncu-rep --page source --csv > source.csv
    python profiles/micro/phase_breakdown.py source.csv [--classes]

For every run of SASS between two BAR.SYNC instructions: SASS count, executed warp instructions (share of the
kernel), stall samples, and the executed instructions per opcode class."""
import csv, sys, collections, re

def cls(op):
    op = op.split('.')[0]
    if op in ('FADD', 'FMUL', 'FFMA'): return 'fp32'
    if op in ('FADD2', 'FMUL2', 'FFMA2'): return 'fp32x2'
    if op in ('MUFU',): return 'mufu'
    if op in ('LDS', 'STS'): return 'smem'
    if op in ('LDG', 'STG'): return 'gmem'
    if op in ('MOV', 'CS2R', 'HFMA2', 'SEL', 'FSEL', 'PRMT'): return 'mov/sel'
    if op in ('FSETP', 'ISETP', 'UISETP', 'PLOP3', 'FMNMX'): return 'cmp'
    if op in ('BRA', 'BSSY', 'BSYNC', 'EXIT', 'CALL', 'RET', 'WARPSYNC'): return 'branch'
    if op in ('BAR',): return 'bar'
    if op in ('I2FP', 'F2I', 'I2F'): return 'cvt'
    if op.startswith('U') or op in ('LDC', 'LDCU', 'S2R', 'S2UR', 'R2UR'): return 'uniform/const'
    return 'int'

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hdr_i]
ci = {n: hdr.index(n) for n in ('Source', '# Samples', 'Instructions Executed')}
phases = [dict(sass=0, inst=0, samp=0, classes=collections.Counter())]
for r in rows[hdr_i + 1:]:
    if len(r) <= ci['Instructions Executed']: continue
    src = r[ci['Source']].strip()
    m = re.match(r'(@!?U?P\w+\s+)?([A-Z0-9_.]+)', src)
    op = m.group(2) if m else '?'
    n = int(float(r[ci['Instructions Executed']] or 0)); s = int(float(r[ci['# Samples']] or 0))
    p = phases[-1]
    p['sass'] += 1; p['inst'] += n; p['samp'] += s; p['classes'][cls(op)] += n
    if op.startswith('BAR'): phases.append(dict(sass=0, inst=0, samp=0, classes=collections.Counter()))
tot = sum(p['inst'] for p in phases); tots = sum(p['samp'] for p in phases)
print("total warp instructions %d, stall samples %d" % (tot, tots))
allc = collections.Counter()
for i, p in enumerate(phases):
    allc.update(p['classes'])
    if p['inst'] < tot * 0.002: continue
    print("phase %2d: sass %4d  inst %9d (%.3f)  samples %5d (%.3f)  %s" % (i, p['sass'], p['inst'], p['inst'] / tot, p['samp'], p['samp'] / max(tots, 1),
          ' '.join('%s:%.1f%%' % (k, 100.0 * v / max(p['inst'], 1)) for k, v in p['classes'].most_common(8))))
print("kernel: " + ' '.join('%s:%.1f%%' % (k, 100.0 * v / tot) for k, v in allc.most_common()))
