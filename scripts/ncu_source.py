"""Aggregate the SASS page of an .ncu-rep into regions: stall samples and executed instructions
per opcode class, and the top-N hottest instructions."""
import csv
import subprocess
import sys
from collections import Counter


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    which = int(sys.argv[3]) if len(sys.argv) > 3 else -1   # kernel index inside the report
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = None
    tables = []
    for r in rows:
        if r and r[0] == 'Kernel Name':
            print('kernel:', r[1][:90])
        if r and r[0] == 'Address':
            hdr = r
            tables.append([])
            continue
        if hdr and len(r) == len(hdr):
            tables[-1].append(r)
    data = tables[which]
    print(f'-- table {which} of {len(tables)}')
    ia, isrc, isamp, iex = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    tot_s = sum(int(r[isamp]) for r in data)
    tot_e = sum(int(r[iex]) for r in data)
    print(f'{len(data)} SASS instructions, {tot_e} executed, {tot_s} samples')
    by_op_e, by_op_s = Counter(), Counter()
    for r in data:
        op = r[isrc].split()[0] if not r[isrc].strip().startswith('@') else r[isrc].split()[1]
        op = op.split('.')[0]
        by_op_e[op] += int(r[iex])
        by_op_s[op] += int(r[isamp])
    print('opcode: executed%  samples%')
    for op, e in by_op_e.most_common(18):
        print(f'  {op:10s} {100 * e / tot_e:5.1f} {100 * by_op_s[op] / max(tot_s, 1):5.1f}')
    print('hottest instructions (index, samples%, executed, main stall, sass):')
    order = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:top]
    for i in sorted(order):
        r = data[i]
        st = max(stall_cols, key=lambda c: int(r[c]))
        print(f'  {i:5d} {100 * int(r[isamp]) / max(tot_s, 1):5.1f} {r[iex]:>8s} {hdr[st]:18s} {r[isrc].strip()[:90]}')
    # cumulative samples by position (10 buckets) to see which phase of the kernel is slow
    n = len(data)
    print('samples by code position decile:', [round(100 * sum(int(r[isamp]) for r in data[n * k // 10:n * (k + 1) // 10]) / max(tot_s, 1), 1) for k in range(10)])
    print('executed by code position decile:', [round(100 * sum(int(r[iex]) for r in data[n * k // 10:n * (k + 1) // 10]) / max(tot_e, 1), 1) for k in range(10)])


if __name__ == '__main__':
    main()
