cd $GRAFT_REPO_ROOT
timeout 300 python tools/dev_timeline.py 60 4 2>&1 | grep "^host: every\|ms_total" | tail -3
python - <<'PY'
import importlib, os, sys, time
sys.path.insert(0, '.')
import torch, bench
synth = importlib.import_module("iridium-sniffer_b200.synth"); pl = importlib.import_module("iridium-sniffer_b200.pipeline")
dev = torch.device("cuda", 0)
iq, _ = bench.make_recording_gpu(torch, synth, 2, 60.0, 100.0, dev); n = iq.shape[0]
p = pl.Pipeline(sample_rate=10_000_000, start_time_ns=10**18)
for i in range(3): p.run_device_raw(iq.data_ptr(), n, "cf32")
torch.cuda.synchronize(); t=time.perf_counter()
for i in range(5): p.run_device_raw(iq.data_ptr(), n, "cf32")
torch.cuda.synchronize(); print("wall per run ms", (time.perf_counter()-t)/5*1e3, "device", p.stats()["ms_total"])
PY
