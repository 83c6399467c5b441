"""Opcode histogram of the tcgen05 / TMEM / bulk-copy instructions per kernel of the in-tree library (cuobjdump -sass; runs without a GPU).

    python tools/sass_summary.py > profiles/r2z_sass_summary.txt
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, 'swem_b200', 'libswem_b200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip().split('(')[0]
COLS = ('UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UBLKCP', 'UBLKRED', 'UTCBAR', 'SYNCS', 'UTMALDG', 'HMMA', 'HGMMA', 'MUFU', 'FMNMX', 'SHFL', 'ELECT')
hist, name = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        name = demangle(m.group(1))
        hist[name] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)', line)
    if m and name:
        hist[name][m.group(1)] += 1
        hist[name]['total'] += 1
print('# cuobjdump -sass swem_b200/libswem_b200.so: opcode histogram of the tcgen05 / TMEM / bulk-copy instructions per kernel (round 2, final)')
print('# UTCHMMA = tcgen05.mma (kind::f16; cta_group::1 and ::2), LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk, UBLKRED = cp.reduce.async.bulk,')
print('# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, ELECT = elect.sync (elected-lane issue); no HMMA / HGMMA (legacy tensor paths) anywhere')
print(f'{"kernel":72s}' + ''.join(f'{c:>8s}' for c in COLS) + f'{"total":>8s}')
for k, h in hist.items():
    if h['UTCHMMA'] or h['UBLKCP'] or 'fusion' in k or 'mkm' in k:
        print(f'{k[:72]:72s}' + ''.join(f'{h[c]:8d}' for c in COLS) + f'{h["total"]:8d}')
