#!/bin/bash
set -u
O=gpurun_out/r2c; mkdir -p $O

echo "== probe"; timeout 300 python scripts/train_probe.py 128 30 2>&1 | tail -3
timeout 300 python scripts/train_probe.py 512 10 2>&1 | tail -1
echo "== ncu launch list of training steps"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 150 --csv --log-file $O/train_launches.csv python scripts/train_probe.py 128 4 > $O/ncu_train.log 2>&1; tail -1 $O/ncu_train.log
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r2c/train_launches.csv')))
hdr = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
cols = rows[hdr]; ki, vi = cols.index('Kernel Name'), cols.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 2:]:
    if len(r) <= vi: continue
    agg[r[ki][:90]][0] += 1; agg[r[ki][:90]][1] += float(r[vi].replace(',', ''))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f'{t / 1e3:10.1f} us  x{n:4d}  {k}')
PY
