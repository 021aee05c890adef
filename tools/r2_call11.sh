#!/bin/bash
# Round 2, GPU call 11 (1 GPU): guided kernel at 2 CTAs / SM with the exact division-free Sampson decision.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -k "guided or golden or pipeline" > gpurun_out/r2c11_pytest.log 2>&1; tail -n 4 gpurun_out/r2c11_pytest.log
timeout 1200 python bench.py --config c5 --images 2000 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2c11_c5_2000.json 2> gpurun_out/r2c11_c5_2000.log
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c11_c5_2000.json").read().strip().splitlines()[-1])
print("c5/2000 N=1", "value", round(d["value"]), "ms/step", round(d["ms_per_step"]), "k1", round(d["k1_ms_per_step"]), "verify", round(d["compact_verify_ms_per_step"]), "pairs", d["config"]["pairs_per_step"], "verified", round(d["config"]["verified_pairs_fraction"],4))
PY
NCU=/usr/local/cuda/bin/ncu
$NCU --metrics gpu__time_duration.sum --clock-control none -k regex:b2m_ -c 3000 --csv --log-file gpurun_out/r2c11_launches_c5_600.csv \
    python bench.py --config c5 --images 600 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2c11_ncu.log 2>&1
python - <<PY
import csv,collections
rows=list(csv.reader(open('gpurun_out/r2c11_launches_c5_600.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
tot=collections.Counter(); cnt=collections.Counter()
for r in rows[start+2:]:
    if len(r)<=vi: continue
    name=r[ki].split('(')[0][-45:]
    try: v=float(r[vi].replace(',',''))
    except: continue
    tot[name]+=v; cnt[name]+=1
s=sum(tot.values())
for k,v in tot.most_common(10): print(f"  {k:45s} n={cnt[k]:4d} total={v/1e6:9.2f} ms share={v/s*100:5.1f}%")
PY
