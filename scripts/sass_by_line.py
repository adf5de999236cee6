"""SASS instruction counts per source line of one kernel (nvdisasm -g -c of a cubin built with -lineinfo).
Usage: python scripts/sass_by_line.py file.cubin kernel-substring source.cu [min_count]"""
import collections, re, subprocess, sys
cubin, kern, srcfile = sys.argv[1:4]
minc = int(sys.argv[4]) if len(sys.argv) > 4 else 10
out = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.split('\n')
cnt = collections.Counter(); cur = None; active = False
for ln in out:
    if ln.startswith('//---') and '.text.' in ln:
        active = kern in ln
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
    elif re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln) and cur:
        cnt[cur] += 1
src = open(srcfile).read().split('\n')
base = srcfile.split('/')[-1]
for (f, l), c in sorted(cnt.items()):
    if f == base and c >= minc:
        print('%5d %5d  %s' % (l, c, src[l - 1].strip()[:120]))
other = collections.Counter()
for (f, l), c in cnt.items():
    if f != base:
        other[f] += c
print('other files:', dict(other), ' total', sum(cnt.values()))
