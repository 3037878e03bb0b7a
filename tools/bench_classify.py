"""Times the frame classification row on one B200: the bench recording (60 s of 10 MHz cf32, ~6.5 k frames) through
the pipeline, then ir_pipeline_classify (device time of k_classify_frames from CUDA events, and wall time of the
whole call including the 504-byte-per-frame copy back) and ir_pipeline_format_parsed_all.  Also classifies a
generated corpus of real IRA / IBC / IDA frames with bit errors (the Chase search does work there; the bench
recording's payloads are random symbols, which fail the access-code / LCW checks early).
Prints one JSON line; a developer aid for profiles/, not part of bench.py's contract.
    python tools/bench_classify.py [seconds]"""
import importlib
import importlib.util
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

synth = importlib.import_module("iridium-sniffer_b200.synth")
pl = importlib.import_module("iridium-sniffer_b200.pipeline")
spec = importlib.util.spec_from_file_location("frame_gen", os.path.join(ROOT, "tests", "frame_gen.py"))
fg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(fg)

dur = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
dev = torch.device("cuda", 0)
iq, _ = bench.make_recording_gpu(torch, synth, 2, dur, 100.0, dev)
n = iq.shape[0]
p = pl.Pipeline(sample_rate=10_000_000, start_time_ns=10**18)
p.run_device_raw(iq.data_ptr(), n, "cf32")
out = {"seconds": dur, "frames": len(p.results().frames)}
ms_k, ms_call, ms_txt = [], [], []
for _ in range(8):
    t = time.perf_counter()
    cls = p.classify()
    ms_call.append((time.perf_counter() - t) * 1e3)
    ms_k.append(p.classify_ms())
    t = time.perf_counter()
    txt = p.parsed_text("T")
    ms_txt.append((time.perf_counter() - t) * 1e3)
out.update(kernel_ms=min(ms_k[2:]), call_ms=min(ms_call[2:]), parsed_all_ms=min(ms_txt[2:]), parsed_bytes=len(txt),
           ida=sum(c.ida_ok for c in cls), ira_ibc=sum(c.frame_type != 0 for c in cls))
p.close()

cases = [c for c in fg.corpus(5, 8000) if c[1] is not None]
t = time.perf_counter()
got = pl.classify_frames(cases)
t1 = time.perf_counter() - t
t = time.perf_counter()
got = pl.classify_frames(cases)
out.update(corpus_frames=len(cases), corpus_call_ms=(time.perf_counter() - t) * 1e3, corpus_first_call_ms=t1 * 1e3,
           corpus_decoded=sum((c.frame_type != 0) + c.ida_ok for c in got))
print(json.dumps(out))
