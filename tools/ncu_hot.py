"""Top stalled SASS instructions of one kernel from an .ncu-rep source page.
Usage: ncu_hot.py <rep> <kernel-regex> [launch-skip] [top-n]"""
import csv
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    skip = sys.argv[3] if len(sys.argv) > 3 else '0'
    topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv',
                          '--kernel-name', f'regex:{pat}', '--launch-skip', skip,
                          '--launch-count', '1'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    col = {n: i for i, n in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) == len(hdr) and r[col['# Samples']].isdigit()]
    tot = sum(int(r[col['# Samples']]) for r in body)
    stall_cols = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
    agg = {n: sum(int(r[col[n]] or 0) for r in body) for n in stall_cols}
    print('total samples', tot)
    print('by reason:', ', '.join(f'{k[6:]}={v}' for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
    ex = sum(int(r[col['Instructions Executed']]) for r in body)
    print('warp instructions executed', ex)
    ranked = sorted(range(len(body)), key=lambda i: -int(body[i][col['# Samples']]))
    for i in ranked[:topn]:
        r = body[i]
        reasons = sorted(((int(r[col[n]] or 0), n[6:]) for n in stall_cols), reverse=True)[:3]
        rs = ' '.join(f'{n}={v}' for v, n in reasons if v)
        print(f'{i:5d} {int(r[col["# Samples"]]):6d} {r[col["Source"]].strip()[:70]:70s} {rs}')
    # shared-memory wavefronts
    wf = sum(int(r[col['L1 Wavefronts Shared']] or 0) for r in body)
    wfi = sum(int(r[col['L1 Wavefronts Shared Ideal']] or 0) for r in body)
    print('shared wavefronts', wf, 'ideal', wfi)


if __name__ == '__main__':
    main()
