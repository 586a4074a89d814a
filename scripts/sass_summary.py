#!/usr/bin/env python
"""Per-kernel SASS opcode summary of libs2vt_b200.so (tcgen05 / TMEM / TMA evidence): python scripts/sass_summary.py > profiles/<name>.md
UTCHMMA = tcgen05.mma (".2CTA": cta_group::2), LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA load; ".MULTICAST" variants),
UTCBAR = tcgen05.commit, SYNCS = mbarrier operations, HMMA = mma.sync (the checker mainloop), UBLKCP = cp.async.bulk (1-D bulk copies: the
peer exchange kernels' loads from / stores to other GPUs)."""
import collections
import os
import re
import subprocess
import sys

lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'multitask-end-to-end-video-captioning_b200', 'libs2vt_b200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
ops = ['UTCHMMA', 'UTCHMMA.2CTA', 'LDTM', 'UTMALDG', 'UTMALDG.MULTICAST', 'UBLKCP', 'UTCBAR', 'SYNCS', 'HMMA', 'UCGABAR', 'REDG']
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r'\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)', line)
    if not m:
        continue
    op = m.group(1)
    c = per[cur]
    if op.startswith('UTCHMMA'):
        c['UTCHMMA'] += 1
        if '2CTA' in op:
            c['UTCHMMA.2CTA'] += 1
    elif op.startswith('LDTM'):
        c['LDTM'] += 1
    elif op.startswith('UTMALDG'):
        c['UTMALDG'] += 1
        if 'MULTICAST' in op:
            c['UTMALDG.MULTICAST'] += 1
    elif op.startswith('UBLKCP'):
        c['UBLKCP'] += 1
    elif op.startswith('UTCBAR'):
        c['UTCBAR'] += 1
    elif op.startswith('SYNCS'):
        c['SYNCS'] += 1
    elif op.startswith('HMMA'):
        c['HMMA'] += 1
    elif op.startswith('UCGABAR'):
        c['UCGABAR'] += 1
    elif op.startswith('REDG') or op.startswith('RED.'):
        c['REDG'] += 1
demangle = subprocess.run(['c++filt'], input='\n'.join(per), capture_output=True, text=True).stdout.splitlines()
tot = collections.Counter()
rows = []
for (name, c), dn in zip(per.items(), demangle):
    if not (c['UTCHMMA'] or c['LDTM'] or c['UTMALDG'] or c['HMMA'] or c['UBLKCP']):
        continue
    dn = re.sub(r'\(.*', '', dn).replace('void ', '')
    rows.append((dn, c))
    tot.update(c)
print('# SASS opcode summary of libs2vt_b200.so (`cuobjdump -sass`, sm_100a)\n')
print('%d kernels in the library, %d of them with tensor-core / TMEM / TMA instructions.  Totals: ' % (len(per), len(rows)) +
      ', '.join('%s %d' % (o, tot[o]) for o in ops if tot[o]) + '\n')
print('| kernel | ' + ' | '.join(ops) + ' |')
print('|---|' + '---:|' * len(ops))
for dn, c in sorted(rows, key=lambda x: -x[1]['UTCHMMA']):
    print('| `%s` | ' % dn[:150] + ' | '.join(str(c[o]) if c[o] else '' for o in ops) + ' |')
