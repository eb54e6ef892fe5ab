"""SASS opcode census of the built objects (build/*.o): per kernel the counts of the Blackwell-specific instructions -
UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAREDG (TMA load / store / reduce), UBLKCP
(bulk copy), UTCBAR (tcgen05.commit), SYNCS (mbarrier), MUFU, and of generic LD / ST (which hot loops must not contain).
usage: python tools/sass_census.py > profiles/r02_sass_census.md"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ['UTCHMMA', 'UTCHMMA.2CTA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMAREDG', 'UBLKCP', 'UTCBAR', 'SYNCS', 'UCGABAR_ARV', 'MUFU', 'REDG',
       'LDS', 'STS', 'LD', 'ST', 'LDL', 'STL']
print('`cuobjdump -sass build/*.o` (nvcc 12.9, `-gencode arch=compute_100a,code=sm_100a`), instruction counts per kernel '
      '(static SASS, not executed counts).  UTCHMMA = tcgen05.mma (UTCHMMA.2CTA = the cta_group::2 form, also counted under UTCHMMA), LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = '
      'TMA tensor load / store / reduce (`.MULTICAST` counted with UTMALDG), UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, '
      'SYNCS = mbarrier ops, UCGABAR = cluster barrier; LD / ST = generic-address loads / stores, LDL / STL = local memory.\n')
print('| object | kernel | total | ' + ' | '.join(OPS) + ' |')
print('|---|---|---:|' + '---:|' * len(OPS))
for obj in sorted(glob.glob(os.path.join(ROOT, 'build', '*.o'))):
    sass = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
    cur, cnt = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            cnt[cur] = collections.Counter()
            continue
        m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', line)
        if m and cur:
            cnt[cur][m.group(1).split('.')[0]] += 1
            if m.group(1).startswith('UTCHMMA.2CTA'):
                cnt[cur]['UTCHMMA.2CTA'] += 1
            cnt[cur]['_total'] += 1
    for k, c in cnt.items():
        name = subprocess.run(['c++filt', k], capture_output=True, text=True).stdout.strip()
        name = re.sub(r'\(.*', '', name).replace('void ', '')
        if c['_total'] < 150 and not any(c[o] for o in OPS[:8]):
            continue
        print(f'| {os.path.basename(obj)} | `{name[:70]}` | {c["_total"]} | ' + ' | '.join(str(c[o]) for o in OPS) + ' |')
