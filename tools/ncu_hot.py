"""Summarise an .ncu-rep: per kernel headline metrics and the hottest SASS instructions by stall samples.
usage: python tools/ncu_hot.py report.ncu-rep [kernel-substring] [top-n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else None
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40


def run(args):
    return subprocess.run(['ncu', '-i', rep] + args, capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(run(['--page', 'raw', '--csv']))))
hdr = raw[0]
want = ['gpu__time_duration.sum', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.avg.per_second']
for r in raw[2:]:
    name = r[hdr.index('Kernel Name')]
    if kern and kern not in name:
        continue
    print('==', name[:60], r[hdr.index('launch__grid_size')] if 'launch__grid_size' in hdr else '')
    for w in want:
        if w in hdr:
            print(f'   {w}: {r[hdr.index(w)]}')
    st = []
    for h, v in zip(hdr, r):
        if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h:
            try:
                st.append((float(v), h.split('issue_stalled_')[1].split('_per_issue')[0]))
            except ValueError:
                pass
    print('   stalls/issue:', ', '.join(f'{n} {v:.2f}' for v, n in sorted(st, reverse=True)[:7]))
if kern:
    src = list(csv.reader(io.StringIO(run(['--page', 'source', '--csv', '--kernel-name', 'regex:' + kern]))))
    # several launches may be concatenated: take the first table
    h = src[1]
    iS, isrc, iex = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed')
    data = []
    for r in src[2:]:
        if len(r) != len(h) or r[0] == 'Address':
            break
        data.append(r)
    tot = sum(int(r[iS]) for r in data)
    print('total samples', tot, 'instructions', len(data))
    sc = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not' not in x]
    top = sorted(enumerate(data), key=lambda x: -int(x[1][iS]))[:topn]
    for i, r in sorted(top):
        s2 = sorted(((int(r[c]), h[c][6:]) for c in sc), reverse=True)[:2]
        print(f'{i:5d} {int(r[iS]):6d} {r[iex]:>8} {r[isrc].strip()[:64]:64} {s2}')
