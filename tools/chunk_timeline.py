import importlib, os, sys, torch
sys.path.insert(0, '/root/repo')
import bench
synth = importlib.import_module("iridium-sniffer_b200.synth")
pl = importlib.import_module("iridium-sniffer_b200.pipeline")
dev = torch.device("cuda", 0)
iq, _ = bench.make_recording_gpu(torch, synth, 2, 60.0, 100.0, dev)
n = iq.shape[0]
host = torch.empty((n, 2), dtype=torch.float32, pin_memory=True); host.copy_(iq); torch.cuda.synchronize()
p = pl.Pipeline(sample_rate=10_000_000, start_time_ns=10**18)
for i in range(3):
    if i == 2: os.environ["IR_CHUNK_DEBUG"] = "1"
    p.run_host_raw(host.data_ptr(), n, "cf32")
p.close()
