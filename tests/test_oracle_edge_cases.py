"""The detector edge cases the GPU tests check against the CPU restatement -- squelch, a burst longer
than max_burst_len (forced baseline update), dense traffic, ragged lengths with bursts inside the
priming period -- pinned here, on the CPU, against the reference's own detector (oracle/_ref,
built with the shared FFT so that every field must agree bit for bit)."""
import numpy as np
import pytest



def _same_bursts(port, ref_dif, iq, min_bursts=0, expect_squelch=None):
    P = port.det_params()
    pb, _, nsq = port.detect(P, iq)
    rb = ref_dif.detect(iq)
    assert len(pb) == len(rb) and len(pb) >= min_bursts, (len(pb), len(rb))
    if expect_squelch is not None:
        assert (nsq > 0) == expect_squelch
    for a, b in zip(rb, pb):
        assert (a["id"], a["start"], a["stop"], a["last_active"], a["center_bin"]) == \
               (b.id, b.start, b.stop, b.last_active, b.center_bin)
        assert a["magnitude"] == b.magnitude and a["noise"] == b.noise
        got = port.extract(P, iq, b)
        assert got.shape == a["samples"].shape and got.tobytes() == a["samples"].tobytes()   # incl. the stale tail
    return pb


def test_squelch_matches_reference(port, ref_dif, synth):
    iq = synth.make_tone_recording(5, 236, 0.02, 0.5)                       # 236 carriers at once > max_bursts = 200
    _same_bursts(port, ref_dif, iq, expect_squelch=True)


def test_too_long_burst_matches_reference(port, ref_dif, synth):
    iq = synth.make_tone_recording(6, 1, 0.13, 0.45, total_s=0.75)          # 130 ms carrier > max_burst_len (90 ms)
    pb = _same_bursts(port, ref_dif, iq, min_bursts=1, expect_squelch=False)
    assert any(b.stop - b.start > 900000 for b in pb)


def test_ragged_length_and_leading_bursts_match_reference(port, ref_dif, synth):
    rec = synth.make_recording(21, duration_s=0.9, n_bursts=10, starts_s=np.linspace(0.05, 0.8, 10))
    _same_bursts(port, ref_dif, rec.iq[:-12345], min_bursts=1)


def test_dense_traffic_matches_reference(port, ref_dif, synth):
    rec = synth.make_dense_recording(1234)               # BASELINE config 4: 672 bursts, ~170 alive at once
    _same_bursts(port, ref_dif, rec.iq, min_bursts=600, expect_squelch=False)


@pytest.mark.parametrize("seed", [101, 102, 103])
def test_random_recordings_match_reference(port, ref_dif, synth, seed):
    rec = synth.make_recording(seed, duration_s=0.75, n_bursts=8)
    _same_bursts(port, ref_dif, rec.iq, min_bursts=3)
