"""One markdown row per profiled launch of an .ncu-rep (`ncu --set full`): duration, grid, registers, tensor / XU / issue
utilisation, DRAM throughput, L2 hit rate, shared-memory pipe, top stall reasons.  usage: python tools/ncu_table.py a.ncu-rep ..."""
import csv
import io
import re
import subprocess
import sys

COLS = [('gpu__time_duration.sum', 'us', 1.0), ('launch__grid_size', 'grid', 1.0), ('launch__registers_per_thread', 'regs', 1.0),
        ('sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', 'tensor %', 1.0),
        ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'XU %', 1.0),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue %', 1.0),
        ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM %', 1.0),
        ('dram__bytes_read.sum', 'DRAM rd MB', 1.0), ('dram__bytes_write.sum', 'DRAM wr MB', 1.0),
        ('lts__t_sector_hit_rate.pct', 'L2 hit %', 1.0),
        ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smem pipe %', 1.0),
        ('sm__cycles_elapsed.avg.per_second', 'SM GHz', 1.0)]
print('| kernel | ' + ' | '.join(c[1] for c in COLS) + ' | top stalls (warp-cycles per issued instruction) |')
print('|---|' + '---:|' * len(COLS) + '---|')
for rep in sys.argv[1:]:
    raw = list(csv.reader(io.StringIO(subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout)))
    hdr, units = raw[0], raw[1]
    for r in raw[2:]:
        name = re.sub(r'\(.*', '', r[hdr.index('Kernel Name')]).replace('void ', '')
        cells = []
        for key, _, _ in COLS:
            if key in hdr:
                v, u = r[hdr.index(key)], units[hdr.index(key)]
                try:
                    x = float(v.replace(',', ''))
                    if u == 'Gbyte':
                        x *= 1e3
                    elif u == 'Kbyte':
                        x /= 1e3
                    elif u == 'byte':
                        x /= 1e6
                    elif u == 'ms':
                        x *= 1e3
                    elif u == 'ns':
                        x /= 1e3
                    cells.append(f'{x:.1f}' if x < 1e4 else f'{x:.0f}')
                except ValueError:
                    cells.append(v)
            else:
                cells.append('-')
        st = []
        for h, v in zip(hdr, r):
            if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h:
                try:
                    st.append((float(v), h.split('issue_stalled_')[1].split('_per_issue')[0]))
                except ValueError:
                    pass
        stalls = ', '.join(f'{n} {v:.2f}' for v, n in sorted(st, reverse=True)[:4])
        print(f'| `{name[:58]}` {r[hdr.index("Grid Size")] if "Grid Size" in hdr else ""} | ' + ' | '.join(cells) + f' | {stalls} |')
