"""Executed warp instructions per CUDA source line of one kernel, from an .ncu-rep captured with --import-source on and the cubin of
the same build (the ncu source page is per SASS instruction; nvdisasm -g gives the line of every instruction offset).
usage: python scripts/ncu_by_line.py report.ncu-rep kernel-regex file.cubin mangled-substring [top]"""
import collections, csv, re, subprocess, sys
rep, kre, cubin, sub = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = next(r for r in rows if 'Instructions Executed' in r)
ia, ie = hdr.index('Address'), hdr.index('Instructions Executed')
data = [r for r in rows if len(r) == len(hdr) and r[ia].startswith('0x')]
base = int(data[0][ia], 16)
out = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.split('\n')
active, cur, line_of = False, None, {}
for ln in out:
    if ln.startswith('//---') and '.text.' in ln:
        active = sub in ln
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
agg, byfile, tot = collections.Counter(), collections.Counter(), 0
for r in data:
    n = int(r[ie] or 0)
    cur = line_of.get(int(r[ia], 16) - base) or ('?', 0)
    agg[cur] += n; byfile[cur[0]] += n; tot += n
src = {}
print('total warp instructions', tot)
for (f, l), n in agg.most_common(top):
    if f not in src:
        try:
            src[f] = open('deepxi_b200/csrc/' + f).read().split('\n')
        except OSError:
            src[f] = None
    text = src[f][l - 1].strip()[:100] if src[f] and 0 < l <= len(src[f]) else ''
    print('%-22s %4d %10d %5.2f%%  %s' % (f, l, n, 100.0 * n / tot, text))
print({k: '%.1f%%' % (100.0 * v / tot) for k, v in byfile.items()})
