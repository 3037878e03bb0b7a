set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_detector_stress.py -x -q 2>&1 | tail -8
timeout 300 python tools/dev_timeline.py 60 3 2>&1 | grep -v "^scan \|^stream scan" | tail -28
timeout 600 ncu -k regex:"k_seg|k_detect_classify|k_detect_scan" --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv python tools/dev_timeline.py 60 1 > gpurun_out/ncu_a.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/r2b_launches.csv') if l.startswith('"')))
h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value")
agg = collections.OrderedDict()
seq = []
for r in rows[1:]:
    k = r[ki].split('(')[0][:60]; v = float(r[vi].replace(',', ''))
    a = agg.setdefault(k, [0, 0.0, 0.0]); a[0] += 1; a[1] += v; a[2] = max(a[2], v)
    seq.append((k[:14], round(v/1e3,1)))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:60s} n={a[0]:4d} sum={a[1]/1e3:10.1f} us max={a[2]/1e3:8.1f} us")
print(seq[:76])
PY
