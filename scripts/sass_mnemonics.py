"""Static SASS mnemonic counts per kernel of the built objects (cuobjdump -sass; runs without a GPU) -> profiles/r02_sass_mnemonics.txt (or the path given).
Evidence that the hot kernels are tcgen05 / TMEM / bulk-copy / mbarrier code and where packed fp32x2 math and local memory appear."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJS = ['tcn_chain.o', 'tcn_umma.o', 'mha_umma.o', 'attn_umma.o', 'stft.o', 'gain.o', 'train_tgt.o', 'tcn_f32.o', 'mhanet.o']
WANT = ['UTCHMMA', 'UTCBAR', 'UTCCP', 'LDTM', 'STTM', 'UTMALDG', 'UTMAPF', 'UBLKCP', 'UBLKPF', 'SYNCS', 'USETMAXREG', 'LDGSTS', 'FFMA2', 'FADD2', 'FMUL2', 'MUFU',
        'ACQBULK', 'ELECT', 'CCTL', 'MEMBAR', 'ATOMG', 'RED', 'LDG', 'STG', 'LDS', 'STS', 'LDL', 'STL']
out = ['SASS mnemonic counts per kernel (cuobjdump -sass of the objects linked into deepxi_b200/libdeepxi_b200.so, sm_100a; static',
       'counts, written by scripts/sass_mnemonics.py).  UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit ->',
       'mbarrier, UTMALDG = cp.async.bulk.tensor (tensor-map TMA), UBLKCP = cp.async.bulk (TMA engine, 1-D), UBLKPF = bulk L2 prefetch, SYNCS = mbarrier operations, USETMAXREG = setmaxnreg,',
       'LDGSTS = cp.async, FFMA2 / FADD2 / FMUL2 = packed fp32x2, LDL / STL = local memory (spills).', '']
for o in OBJS:
    path = os.path.join(ROOT, 'deepxi_b200', 'build', o)
    if not os.path.exists(path):
        continue
    txt = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True).stdout
    cur, counts = None, collections.OrderedDict()
    for ln in txt.splitlines():
        m = re.search(r'Function : (\S+)', ln)
        if m:
            cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip().split('(')[0]
            counts[cur] = collections.Counter()
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', ln)
        if m and cur:
            counts[cur][m.group(1)] += 1
            counts[cur]['_total'] += 1
    for k, c in counts.items():
        if c['_total'] < 50:
            continue
        out.append('%s  [%s, %d instructions]' % (k, o, c['_total']))
        out.append('    ' + ', '.join('%s %d' % (w, c[w]) for w in WANT if c[w]))
dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'r02_sass_mnemonics.txt')
open(dst, 'w').write('\n'.join(out) + '\n')
print('wrote', dst, len(out), 'lines')
