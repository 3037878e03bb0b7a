"""Developer tool: run the bench recording once with IR_SCAN_DEBUG=1 so that the cluster state
machine prints its per-phase cycle counters (leader CTA and one owner CTA)."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["IR_SCAN_DEBUG"] = "1"

import torch  # noqa: E402

import bench  # noqa: E402

synth = importlib.import_module("iridium-sniffer_b200.synth")
pl = importlib.import_module("iridium-sniffer_b200.pipeline")
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
dev = torch.device("cuda", 0)
iq, truth = bench.make_recording_gpu(torch, synth, 2, secs, 100.0, dev)
p = pl.Pipeline(sample_rate=10_000_000, start_time_ns=1)
for i in range(2):
    r = p.run_device_ptr(iq.data_ptr(), iq.shape[0], "cf32")
    print("scan ms", r.stats["ms_detect_scan"], "bursts", len(r.bursts), "frames", iq.shape[0] // 8192)
