"""Debug aid: run the bench recording through the streaming state machine and through the cluster
kernel (IR_SCAN=cluster) and report the first burst on which they differ."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
synth = importlib.import_module("iridium-sniffer_b200.synth")
pl = importlib.import_module("iridium-sniffer_b200.pipeline")
dev = torch.device("cuda", 0)
iq, _ = bench.make_recording_gpu(torch, synth, 2, secs, 100.0, dev)
n = iq.shape[0]
KEYS = ("id", "start", "stop", "last_active", "center_bin", "magnitude", "noise")


def run(mode):
    if mode:
        os.environ["IR_SCAN"] = mode
    else:
        os.environ.pop("IR_SCAN", None)
    p = pl.Pipeline(sample_rate=10_000_000, start_time_ns=10**18)
    r = p.run_device_ptr(iq.data_ptr(), n, "cf32")
    ss = p.scan_stats()
    p.close()
    return [tuple(b[k] for k in KEYS) for b in r.bursts], ss


a, sa = run(None)
b, sb = run("cluster")
print("stream", len(a), sa)
print("cluster", len(b))
for i, (x, y) in enumerate(zip(a, b)):
    if x != y:
        print("first difference at burst", i)
        for j in range(max(0, i - 3), min(len(a), len(b), i + 6)):
            print(" s", a[j], "\n c", b[j], "" if a[j] == b[j] else "   <-- differs")
        break
else:
    print("identical over the common prefix")
