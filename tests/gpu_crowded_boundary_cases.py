"""A chunk boundary INSIDE a crowded stretch: the dense recording (672 bursts in 150 ms) behind 2.5 M samples of noise,
copied in 2 Mi-sample pieces which the host joins into chunks of 4 Mi samples -- the second boundary (sample 8 388 608)
falls where 112 bursts are alive.  The list longer than a warp leaves one chunk through k_seg_commit (DetState.act),
enters the next through k_seg_begin and the overflow arrays, and its first segment goes straight to the plain walker.
The oracle's burst list and frame bits, no chunk handed to the cluster kernel.

Written after the round's last GPU minute (the pieces it is made of ran: tests/test_gpu_detector_stress.py
::test_dense_traffic_fed_in_pieces, with the crowded stretch inside one chunk); tests/test_zz_gpu_classify.py runs it
last, in a child process."""
import importlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_crowded_chunk_boundary(port, synth):
    pl = importlib.import_module("iridium-sniffer_b200.pipeline")
    dense = synth.make_dense_recording(1234)
    rng = np.random.default_rng(77)
    lead = (rng.standard_normal(2_500_000) + 1j * rng.standard_normal(2_500_000)).astype(np.complex64) * np.float32(0.01)
    iq = np.concatenate([lead, dense.iq])
    P = port.det_params()
    pb, _, nsq = port.detect(P, iq)
    want = [(o.id, o.start, o.stop, o.last_active, o.center_bin, o.magnitude, o.noise) for o in pb]
    cut = 8 << 20
    assert nsq == 0 and len(want) > 600
    assert sum(1 for b in pb if b.start + P.burst_pre_len <= cut < b.stop) > 64        # alive at the boundary
    os.environ.pop("IR_SCAN", None)                                                     # the segmented state machine
    p = pl.Pipeline(sample_rate=10_000_000, start_time_ns=77, h2d_chunk=2 << 20)
    res = p.run_host(iq, "cf32")
    ss = p.scan_stats()
    p.close()
    got = [(b["id"], b["start"], b["stop"], b["last_active"], b["center_bin"], b["magnitude"], b["noise"]) for b in res.bursts]
    assert got == want, ss
    assert ss["segmented"] and ss["launches_bailed"] == 0 and ss["generic_segment_walks"] > 0 and ss["launches_kept"] >= 3, ss
    ores, _ = port.run(iq, start_time_ns=77)
    assert [(f["id"], f["bits"].tobytes()) for f in res.frames] == [(o["id"], o["bits"].tobytes()) for o in ores]
