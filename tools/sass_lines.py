"""Per-source-line SASS instruction counts of a kernel between two BAR.SYNCs
or within an address window.  Usage:
  sass_lines.py <obj> <kernel-substr> [lo hi]   (lo/hi: hex instruction offsets)
Needs -lineinfo.  Prints the loop windows if no window is given."""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else None
    hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else None
    d = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=d,
                   capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith('.cubin')][0]
    dis = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(d, cub)],
                         capture_output=True, text=True).stdout.split('\n')
    inside = False
    cur = ('?', 0)
    per_line = collections.Counter()
    ops = collections.defaultdict(collections.Counter)
    n = 0
    for l in dis:
        if l.startswith('//---') and '.text.' in l:
            inside = pat in l
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m:
            a = int(m.group(1), 16)
            if (lo is None or a >= lo) and (hi is None or a <= hi):
                t = m.group(2).split()
                op = t[1] if t[0].startswith('@') else t[0]
                per_line[cur] += 1
                ops[cur][op.split('.')[0]] += 1
                n += 1
    print('instructions in window:', n)
    for (f, ln), c in sorted(per_line.items(), key=lambda kv: -kv[1])[:70]:
        top = ' '.join(f'{k}:{v}' for k, v in ops[(f, ln)].most_common(6))
        print(f'{c:5d}  {f}:{ln:<5d} {top}')


if __name__ == '__main__':
    main()
