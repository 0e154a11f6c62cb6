"""List the backward branches (loops) of a kernel's SASS with their body size
and opcode histogram.  Usage: sass_loops.py <obj> <mangled-name-substring>"""
import collections
import re
import subprocess
import sys


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    names = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True,
                           text=True).stdout
    blocks = re.split(r'\n\s*Function : ', names)
    for blk in blocks[1:]:
        name = blk.split('\n', 1)[0].strip()
        if pat not in name:
            continue
        ins = []
        for l in blk.split('\n'):
            m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        print(name, len(ins), 'instructions')
        addr = {a: i for i, (a, _) in enumerate(ins)}
        for i, (a, t) in enumerate(ins):
            m = re.search(r'BRA\S*\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)', t)
            if m:
                tgt = int(m.group(1), 16)
                if tgt < a and tgt in addr:
                    body = ins[addr[tgt]:i + 1]
                    if len(body) < 200:
                        continue
                    h = collections.Counter()
                    for _, tt in body:
                        op = tt.split()[0]
                        if op.startswith('@'):
                            op = tt.split()[1]
                        h[op.split('.')[0]] += 1
                    print(f'loop {hex(tgt)}..{hex(a)}: {len(body)} instr')
                    print('  ', ', '.join(f'{k}:{v}' for k, v in h.most_common(40)))


if __name__ == '__main__':
    main()
