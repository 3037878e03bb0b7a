"""The FIR kernel's shared memory starts as whatever the previous kernel on that SM left there.  A chain of the
register-tiled FIR multiplies a few samples past an output's last input by ZERO taps (the zero-padded last tap block at
DEC = 48, the unpeeled middle blocks for the last outputs of a burst's last tile); if such a sample is a NaN bit pattern
-- another kernel's bitmap words are -- the sum is poisoned.  It happened: ~4 % of a 12 MHz recording's frames were lost
until the staging wrote zeros there.  IR_FIR_POISON=1 makes every FIR CTA fill its sample buffers with NaNs before it
starts, so that any such read shows up as a frame that differs from the oracle's.

The switch is read at the first FIR launch of a process: tests/test_zz_gpu_classify.py runs this file in a child."""
import importlib
import os

import pytest

os.environ["IR_FIR_POISON"] = "1"

pytestmark = pytest.mark.gpu

FS12 = 12_000_000


def _frames_equal(pl, port, iq, fs, chunk=0):
    p = pl.Pipeline(sample_rate=fs, start_time_ns=77, h2d_chunk=chunk)
    res = p.run_host(iq, "cf32")
    p.close()
    ores, _ = port.run(iq, sample_rate=fs, start_time_ns=77)
    got = [(f["id"], f["bits"].tobytes()) for f in res.frames]
    want = [(o["id"], o["bits"].tobytes()) for o in ores]
    assert len(want) >= 8
    assert got == want
    return len(want)


def test_fir_ignores_what_shared_memory_held(port, synth):
    pl = importlib.import_module("iridium-sniffer_b200.pipeline")
    # 12 MHz (DEC = 48, 240-output tiles): bursts of every length, many small waves
    rec = synth.make_recording(34, sample_rate=FS12, duration_s=1.7, n_bursts=20)
    _frames_equal(pl, port, rec.iq, FS12, chunk=2 << 20)
    # 10 MHz (DEC = 40): last tiles of every fill
    rec = synth.make_recording(102, duration_s=1.2, n_bursts=16)
    _frames_equal(pl, port, rec.iq, 10_000_000)
