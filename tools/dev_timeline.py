"""Developer tool: device-resident run of the bench recording with the per-chunk / per-wave timeline
(IR_CHUNK_DEBUG) and the state machine's counters (IR_SCAN_DEBUG) of the last of three runs."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
synth = importlib.import_module("iridium-sniffer_b200.synth")
pl = importlib.import_module("iridium-sniffer_b200.pipeline")
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
skip_s = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0      # start the run this far into the recording (a time block)
dev = torch.device("cuda", 0)
iq, _ = bench.make_recording_gpu(torch, synth, 2, secs, 100.0, dev)
off = int(skip_s * 10_000_000) // 32768 * 32768
n = iq.shape[0] - off
p = pl.Pipeline(sample_rate=10_000_000, start_time_ns=10**18)
for i in range(runs):
    if i == runs - 1:
        os.environ["IR_CHUNK_DEBUG"] = "1"
        os.environ["IR_SCAN_DEBUG"] = "1"
    p.run_device_raw(iq.data_ptr() + off * 8, n, "cf32")
    st = p.stats()
    print({k: round(v, 3) if isinstance(v, float) else v for k, v in st.items()})
import json
r = p.results()
fs = 10_000_000
print("RUNJSON " + json.dumps({"det_frames": n // 8192, "bursts": len(r.bursts),
      "tiles": sum((b["dec_len"] + 255) // 256 for b in r.bursts if b["dec_len"] >= 100), "samples": n}))
p.close()
