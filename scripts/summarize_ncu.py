#!/usr/bin/env python
"""Turn an .ncu-rep (brought back in gpurun_out/) into small text summaries under profiles/.

    python scripts/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/ncu_r1_xxx
writes <prefix>_metrics.csv (selected raw metrics per captured launch) and <prefix>_hotlines.txt
(warp-stall samples per CUDA source line for each distinct kernel; needs -lineinfo + --import-source on)."""
import csv
import io
import subprocess
import sys

KEYS = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max']


def run(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def main(rep, prefix):
    rows = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'raw', '--csv']))))
    h = rows[0]
    idx = [h.index(k) for k in KEYS if k in h]
    with open(prefix + '_metrics.csv', 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow([h[i] for i in idx]); w.writerow([rows[1][i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
    # per-kernel DRAM traffic (bytes per launch, mean over the captured launches) for bench.py's roofline.traffic
    import json, os, collections
    tr = collections.defaultdict(list)
    ni, ri, wi = h.index('Kernel Name'), h.index('dram__bytes_read.sum'), h.index('dram__bytes_write.sum')
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    for r in rows[2:]:
        short = r[ni].split('(')[0].split('<')[0].split('::')[-1].split()[-1]
        tr[short].append(float(r[ri]) * scale[rows[1][ri]] + float(r[wi]) * scale[rows[1][wi]])
    tj = os.path.join(os.path.dirname(prefix), 'ncu_traffic.json')
    old = json.load(open(tj)) if os.path.exists(tj) else {}
    old.update({k: sum(v) / len(v) for k, v in tr.items()})
    old['_source'] = os.path.basename(prefix) + '_metrics.csv (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)'
    json.dump(old, open(tj, 'w'), indent=1)
    names = []
    for r in rows[2:]:
        n = r[h.index('Kernel Name')]
        if n not in names:
            names.append(n)
    with open(prefix + '_hotlines.txt', 'w') as f:
        for n in names:
            short = n.split('(')[0].split('<')[0].split('::')[-1].split()[-1]
            out = run(['-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv', '--kernel-name', 'regex:' + short,
                       '--launch-count', '1'])
            cur, lines = None, []
            for r in csv.reader(io.StringIO(out)):
                if r and r[0] == 'File Path':
                    cur = r[1].split('/')[-1]
                elif len(r) > 5 and r[0] not in ('', 'Line No', 'Function Name') and r[2] == '-':
                    try:
                        lines.append((int(r[4]), cur, int(r[0]), r[1].strip()[:110]))
                    except ValueError:
                        pass
            tot = sum(x[0] for x in lines) or 1
            f.write(f'=== {n}\n    warp-stall samples: {tot}\n')
            for s, fl, ln, src in sorted(lines, reverse=True)[:18]:
                f.write(f'{s:7d} {100 * s / tot:5.1f}%  {fl}:{ln}  {src}\n')
            f.write('\n')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
