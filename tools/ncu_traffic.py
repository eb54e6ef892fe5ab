"""ncu launch list (gpu__time_duration + dram__bytes_read/write, one profiled step of `bench.py --profile-step`) ->
per-family table (markdown) and the `roofline.traffic` entry of profiles/traffic.json.
Usage: python tools/ncu_traffic.py launches.csv BATCH LATENT out.md"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MULT = {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9, 'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def main(path, batch, latent, out_md):
    rows = collections.defaultdict(dict)
    with open(path, newline='') as f:
        lines = [l for l in f if not l.startswith('==')]
    for r in csv.DictReader(lines):
        v = float(r['Metric Value'].replace(',', '')) * MULT.get(r['Metric Unit'], 1)
        rows[r['ID']][r['Metric Name']] = v
        rows[r['ID']]['name'] = re.sub(r'\(.*', '', r['Kernel Name'])
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for d in rows.values():
        a = agg[d['name']]
        a[0] += 1
        a[1] += d.get('gpu__time_duration.sum', 0)
        a[2] += d.get('dram__bytes_read.sum', 0)
        a[3] += d.get('dram__bytes_write.sum', 0)
    tot_ns = sum(a[1] for a in agg.values())
    out = [f'`ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum '
           f'--clock-control none python bench.py --profile-step --batch {batch} --latent {latent}`: one warmed-up step, '
           f'{sum(a[0] for a in agg.values())} launches, {tot_ns / 1e6:.2f} ms of kernel time (serialised under ncu)', '',
           '| kernel | launches | ms | share | avg us | DRAM read GB | DRAM write GB | DRAM GB/s |', '|---|---:|---:|---:|---:|---:|---:|---:|']
    for n, (c, t, r, w) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f'| `{n[:90]}` | {c} | {t / 1e6:.3f} | {100 * t / tot_ns:.1f}% | {t / c / 1e3:.1f} | {r / 1e9:.2f} | {w / 1e9:.2f} | {(r + w) / max(t, 1):.0f} |')
    fam = [a for n, a in agg.items() if 'gemm_tc' in n or 'attn_fwd' in n or 'attn_bwd' in n]
    fam_bytes = sum(a[2] + a[3] for a in fam)
    out += ['', f'Tensor-core family (gemm_tc_kernel + attn_fwd/bwd_kernel): {sum(a[0] for a in fam)} launches, '
            f'{sum(a[1] for a in fam) / 1e6:.2f} ms ({100 * sum(a[1] for a in fam) / tot_ns:.1f}% of the step), '
            f'{fam_bytes / 1e9:.2f} GB of DRAM traffic per step.']
    text = '\n'.join(out) + '\n'
    print(text)
    with open(out_md, 'w') as f:
        f.write(text)
    tj = os.path.join(ROOT, 'profiles', 'traffic.json')
    table = json.load(open(tj)) if os.path.exists(tj) else []
    table = [r for r in table if not (r['per_gpu_microbatch'] == batch and r['latent'] == latent)]
    table.append({'per_gpu_microbatch': batch, 'latent': latent, 'tensor_family_dram_bytes_per_step': fam_bytes,
                  'tensor_family_ms_under_ncu': sum(a[1] for a in fam) / 1e6, 'source': os.path.basename(out_md)})
    with open(tj, 'w') as f:
        json.dump(table, f, indent=1)


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
