"""Write SASS listings of the hot loops / kernels to profiles/ (evidence for
DESIGN.md: north_star asks for SASS listings next to the counters).

    python tools/sass_dump.py            # after the library has been built
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import sass_model as sm  # noqa: E402

BUILD = os.path.join(ROOT, 'uncertainty_model_b200', 'build')
OUT = os.path.join(ROOT, 'profiles')


def listing(obj, pat):
    out = subprocess.run(['cuobjdump', '-sass', os.path.join(BUILD, obj)],
                         capture_output=True, text=True).stdout
    for blk in re.split(r'\n\s*Function : ', out)[1:]:
        if pat in blk.split('\n', 1)[0]:
            name = blk.split('\n', 1)[0].strip()
            lines = [re.sub(r'\s*/\* 0x[0-9a-f]+ \*/\s*$', '', l).rstrip()
                     for l in blk.split('\n')
                     if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l)]
            return name, lines
    raise SystemExit(f'{pat} not found in {obj}')


def loops(obj, pat):
    ins = sm.parse(os.path.join(BUILD, obj), pat)
    addr = {a: k for k, (a, _, _) in enumerate(ins)}
    found = []
    for k, (a, t, _) in enumerate(ins):
        m = re.search(r'BRA\S*\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)', t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr:
                found.append((tgt, a, k - addr[tgt] + 1))
    return ins, found


def dump(fname, obj, pat, pick, title):
    name, lines = listing(obj, pat)
    ins, found = loops(obj, pat)
    lo, hi, n = pick(found)
    body = [l for l in lines
            if lo <= int(re.match(r'\s+/\*([0-9a-f]+)\*/', l).group(1), 16) <= hi]
    T, nn, count, _, wait = sm.model(ins, lo, hi)
    with open(os.path.join(OUT, fname), 'w') as f:
        f.write(f'# {title}\n# {name}\n# loop {hex(lo)}..{hex(hi)}: {n} instructions; '
                f'lone-warp issue model {T} cycles\n'
                f'# opcodes: {", ".join(f"{k}:{v}" for k, v in count.most_common(30))}\n')
        f.write('\n'.join(body) + '\n')
    print(fname, n, 'instructions')


def whole(fname, obj, pat, title):
    name, lines = listing(obj, pat)
    with open(os.path.join(OUT, fname), 'w') as f:
        f.write(f'# {title}\n# {name}\n# {len(lines)} instructions\n')
        f.write('\n'.join(lines) + '\n')
    print(fname, len(lines), 'instructions')


def main():
    # steady-state loop of the fused kernel: two column steps, the middle one of
    # the three big loops
    dump('r02_sass_col_kernel_steady_loop.txt', 'col_inst_512.o', 'ILi528ELb1ELi0ELi47',
         lambda f: sorted([x for x in f if 1000 < x[2] < 2000])[1],
         'col_kernel<528,GRAD,PLAIN,hot terms>: steady-state loop (two rows per trip)')
    ins_r, _ = loops('cons_rows.o', 'cons_rows_kernel')

    def walk(found):        # the loop of scalar read-modify-writes: no global
        best = None         # loads, no reductions, >= 32 shared stores
        for lo, hi, n in found:
            _, _, count, _, _ = sm.model(ins_r, lo, hi)
            if count['LDG'] == 0 and count['REDG'] == 0 and count['LDGSTS'] == 0 \
                    and count['STS'] >= 32 and count['FADD'] >= 32 \
                    and (best is None or n < best[2]):
                best = (lo, hi, n)
        return best
    dump('r02_sass_cons_rows_walk_loop.txt', 'cons_rows.o', 'cons_rows_kernel', walk,
         'cons_rows_kernel: the walk (batches of 8 columns, software pipelined)')
    whole('r02_sass_cons_scatter2.txt', 'cons_kernels.o', 'cons_scatter2_kernelILb1',
          'cons_scatter2_kernel<aligned>: warp-per-row scatter (stand-alone terms)')
    whole('r02_sass_spars_scatter.txt', 'spars.o', 'scatter_kernelILi1',
          'spars scatter_kernel<values>: one radix pass of a tile')
    whole('r02_sass_spars_pool11.txt', 'spars.o', 'pool_kernelILi11',
          'spars pool_kernel<11>: exact 11x11 mean + key')


if __name__ == '__main__':
    main()
