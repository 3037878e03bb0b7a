"""The streaming state machine's ALGORITHM (tests/model_stream_scan.py: guard-banded bitmaps, deadlines in
frames, bursts ordered by id, events only where the bitmaps do not prove the outcome) reproduces the CPU
oracle's burst list field for field, launch boundaries included, and gives up exactly on the cases the
kernel hands to the cluster kernel.  No GPU."""
import importlib.util
import os

import numpy as np
import pytest

_spec = importlib.util.spec_from_file_location("model_stream_scan", os.path.join(os.path.dirname(__file__), "model_stream_scan.py"))
msm = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(msm)


def _model(P):
    return msm.StreamScanModel(P.fft_size, P.threshold_lin, P.burst_width_bins // 2, P.burst_pre_len, P.burst_post_len,
                               P.max_burst_len, P.max_bursts, P.history_size)


def _oracle(port, iq):
    P = port.det_params()
    pb, mag, nsq = port.detect(P, iq, dump_mag=True)
    want = [(b.id, b.start, b.stop, b.last_active, b.center_bin, b.peak_rel, b.base_at_create) for b in pb]
    return P, mag, want, nsq


@pytest.mark.parametrize("launch_frames", [700, 128, 4096])
def test_model_equals_oracle_config1(port, rec_small, launch_frames):
    P, mag, want, nsq = _oracle(port, rec_small.iq)
    m = _model(P)
    got = m.run(mag, launch_frames=launch_frames)
    assert nsq == 0 and len(want) == 13
    assert [tuple(g[:5]) for g in got] == [w[:5] for w in want]
    assert all(np.float32(g[5]) == np.float32(w[5]) and np.float32(g[6]) == np.float32(w[6]) for g, w in zip(got, want))
    assert m.stats["events"] < mag.shape[0] // 4           # most frames are decided by the bitmaps alone


@pytest.mark.parametrize("seed", [101, 102])
def test_model_equals_oracle_random(port, synth, seed):
    rec = synth.make_recording(seed, duration_s=0.75, n_bursts=8)
    P, mag, want, _ = _oracle(port, rec.iq)
    got = _model(P).run(mag, launch_frames=100)            # bursts alive across many launch boundaries
    assert len(want) >= 3 and [tuple(g[:5]) for g in got] == [w[:5] for w in want]
    assert all(np.float32(g[5]) == np.float32(w[5]) and np.float32(g[6]) == np.float32(w[6]) for g, w in zip(got, want))


def test_model_gives_up_where_the_kernel_does(port, synth):
    cases = [("squelch", synth.make_tone_recording(5, 236, 0.02, 0.5)),
             ("too long", synth.make_tone_recording(6, 1, 0.13, 0.45, total_s=0.75))]
    for reason, iq in cases:
        P, mag, _, _ = _oracle(port, iq)
        with pytest.raises(msm.Bail) as e:
            _model(P).run(mag)
        assert e.value.reason in (reason, "a 33rd concurrent burst"), e.value.reason


def test_model_guard_band_catches_a_moving_noise_floor(port, synth):
    """bursts inside the priming period inflate the baseline; when they rotate out of the 512-frame
    history the baseline leaves [0.65, 1.5] x reference and the launch must be abandoned, not trusted"""
    rec = synth.make_recording(21, duration_s=0.9, n_bursts=10, starts_s=np.linspace(0.05, 0.8, 10))
    P, mag, want, _ = _oracle(port, rec.iq[:-12345])
    m = _model(P)
    try:
        got = m.run(mag, launch_frames=4096)
    except msm.Bail as e:
        assert e.reason == "guard band"
    else:
        assert [tuple(g[:5]) for g in got] == [w[:5] for w in want]
