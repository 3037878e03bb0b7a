"""The segmented state machine's ALGORITHM (tests/model_seg_scan.py: fixed cuts, speculative burst lists at
the cuts, baseline versions from the previous round's quiet flags, fixed point = exact) reproduces the CPU
oracle's burst list field for field for every segment length, and gives up exactly where the kernels hand the
chunk to the cluster kernel.  No GPU."""
import importlib.util
import os

import numpy as np
import pytest

_spec = importlib.util.spec_from_file_location("model_seg_scan", os.path.join(os.path.dirname(__file__), "model_seg_scan.py"))
msm = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(msm)


def _model(P, seg, **kw):
    return msm.SegScanModel(P.fft_size, P.threshold_lin, P.burst_width_bins // 2, P.burst_pre_len, P.burst_post_len,
                            P.max_burst_len, P.max_bursts, P.history_size, seg=seg, **kw)


@pytest.fixture(scope="module")
def c_walker(tmp_path_factory):
    """the product's generic segment walker (csrc/seg_generic.cuh) compiled for the host"""
    import ctypes as C
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    out = str(tmp_path_factory.mktemp("segg") / "libsegg_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-pthread", "-I/usr/local/cuda/include", "-x", "c++",
                    os.path.join(here, "seg_generic_host_shim.cpp"), "-o", out], check=True)
    lib = C.CDLL(out)
    lib.segg_walk.restype = C.c_int
    lib.segg_walk.argtypes = [C.c_int] * 6 + [C.c_float] + [C.c_int] * 3 + [C.c_longlong, C.c_int] + [C.c_void_p] * 6 + \
                             [C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_uint32), C.c_int]
    return lib


def _oracle(port, iq):
    P = port.det_params()
    pb, mag, nsq = port.detect(P, iq, dump_mag=True)
    want = [(b.id, b.start, b.stop, b.last_active, b.center_bin, b.peak_rel, b.base_at_create) for b in pb]
    return P, mag, want, nsq


def _same(got, want):
    assert [tuple(g[:5]) for g in got] == [w[:5] for w in want]
    assert all(np.float32(g[5]) == np.float32(w[5]) and np.float32(g[6]) == np.float32(w[6]) for g, w in zip(got, want))


@pytest.mark.parametrize("seg,chunk", [(64, 4096), (16, 4096), (7, 300), (256, 700), (4096, 4096)])
def test_model_equals_oracle_config1(port, rec_small, seg, chunk):
    P, mag, want, nsq = _oracle(port, rec_small.iq)
    m = _model(P, seg)
    got = m.run(mag, chunk_frames=chunk)
    assert nsq == 0 and len(want) == 13
    _same(got, want)
    # the speculation settles in a handful of rounds once a segment is longer than a burst lives
    # (burst + post_len ~ 30 frames); shorter segments need a round per segment a burst spans
    assert m.stats["max_rounds_seen"] <= (12 if seg < 32 else 6), m.stats


@pytest.mark.parametrize("seed", [101, 102])
def test_model_equals_oracle_random(port, synth, seed):
    rec = synth.make_recording(seed, duration_s=0.75, n_bursts=8)
    P, mag, want, _ = _oracle(port, rec.iq)
    m = _model(P, 32)
    got = m.run(mag, chunk_frames=100)            # bursts alive across many chunk boundaries
    assert len(want) >= 3
    _same(got, want)


def test_model_gives_up_where_the_kernels_do(port, synth):
    cases = [("squelch", synth.make_tone_recording(5, 236, 0.02, 0.5)),
             ("too long", synth.make_tone_recording(6, 1, 0.13, 0.45, total_s=0.75))]
    for reason, iq in cases:
        P, mag, _, _ = _oracle(port, iq)
        with pytest.raises(msm.Bail) as e:
            _model(P, 64).run(mag)
        assert e.value.reason in (reason, "a 33rd concurrent burst", "more bursts than lanes"), e.value.reason


def test_model_follows_a_moving_noise_floor(port, synth):
    """bursts inside the priming period inflate the baseline; when they rotate out of the 512-frame history the
    baseline leaves [0.65, 1.5] x reference: the bin's band is widened to what the baseline really did, the
    bitmaps are rebuilt, and the result is still the oracle's, field for field"""
    rec = synth.make_recording(21, duration_s=0.9, n_bursts=10, starts_s=np.linspace(0.05, 0.8, 10))
    P, mag, want, _ = _oracle(port, rec.iq[:-12345])
    m = _model(P, 64)
    got = m.run(mag, chunk_frames=4096)
    _same(got, want)
    assert m.stats.get("rebuilds", 0) >= 1, m.stats


# ---------------------------------------------------------------- the generic walker of the kernels, on the CPU
def test_generic_walker_equals_oracle(port, rec_small, synth, c_walker):
    """csrc/seg_generic.cuh -- what k_seg_walk hands a segment with more than 32 bursts to -- walking EVERY segment
    of the model's rounds: the oracle's burst list, field for field"""
    seg = c_walker.segg_seg_len()
    P, mag, want, _ = _oracle(port, rec_small.iq)
    m = _model(P, seg, c_walker=c_walker, lanes=512)
    _same(m.run(mag, chunk_frames=4096), want)
    rec = synth.make_recording(102, duration_s=0.75, n_bursts=8)
    P, mag, want, _ = _oracle(port, rec.iq)
    _same(_model(P, seg, c_walker=c_walker, lanes=512).run(mag, chunk_frames=256), want)
    rec = synth.make_recording(21, duration_s=0.9, n_bursts=10, starts_s=np.linspace(0.05, 0.8, 10))   # moving noise floor
    P, mag, want, _ = _oracle(port, rec.iq[:-12345])
    _same(_model(P, seg, c_walker=c_walker, lanes=512).run(mag, chunk_frames=4096), want)


def test_generic_walker_on_dense_traffic(port, synth, c_walker):
    """BASELINE config 4: ~170 bursts alive at once, far beyond the 32 a fast walker holds -- no bail, the oracle's list"""
    rec = synth.make_dense_recording(1234)
    P, mag, want, nsq = _oracle(port, rec.iq)
    assert nsq == 0 and len(want) >= 600
    m = _model(P, c_walker.segg_seg_len(), c_walker=c_walker, lanes=512)
    _same(m.run(mag, chunk_frames=4096), want)
    assert m.stats["max_rounds_seen"] <= 10, m.stats


@pytest.mark.parametrize("walker_lanes", [3, 4])
def test_generic_walker_with_several_lanes(port, rec_small, synth, c_walker, walker_lanes):
    """the same walker with its lane loops, ballots, scans and reductions really spread over several lanes (threads
    meeting at barriers stand in for the warp): dense traffic and an ordinary recording, the oracle's list"""
    rec = synth.make_dense_recording(1234)
    P, mag, want, _ = _oracle(port, rec.iq)
    m = _model(P, c_walker.segg_seg_len(), c_walker=c_walker, lanes=512, walker_lanes=walker_lanes)
    _same(m.run(mag, chunk_frames=4096), want)
    P, mag, want, _ = _oracle(port, rec_small.iq)
    _same(_model(P, c_walker.segg_seg_len(), c_walker=c_walker, lanes=512, walker_lanes=walker_lanes).run(mag, chunk_frames=4096), want)
    for reason, iq in (("squelch", synth.make_tone_recording(5, 236, 0.02, 0.5)),
                       ("too long", synth.make_tone_recording(6, 1, 0.13, 0.45, total_s=0.75))):
        P, mag, _, _ = _oracle(port, iq)
        with pytest.raises(msm.Bail) as e:
            _model(P, c_walker.segg_seg_len(), c_walker=c_walker, lanes=512, walker_lanes=walker_lanes).run(mag)
        assert e.value.reason == reason, e.value.reason


@pytest.mark.parametrize("walker_lanes", [1, 4])
def test_generic_walker_12mhz_clipped_int16(port, synth, c_walker, walker_lanes):
    """BASELINE config 3's geometry and its trouble in small: 12 MHz, 16384-pt frames, int16 samples of which the
    reference keeps the upper byte; strong bursts clip and every one of them spawns dozens of spurious detections
    (229 bursts from 14 planted, 76 alive at once) -- the segments the device hands to this walker"""
    fs = 12_000_000
    rec = synth.make_recording(3, sample_rate=fs, duration_s=1.6, n_bursts=14, fmt="ci16", snr_db=(26.0, 30.0))
    iq = port.convert_ci16(rec.iq)
    P = port.det_params(sample_rate=fs)
    pb, mag, nsq = port.detect(P, iq, dump_mag=True)
    want = [(b.id, b.start, b.stop, b.last_active, b.center_bin, b.peak_rel, b.base_at_create) for b in pb]
    assert P.fft_size == 16384 and len(want) > 150 and nsq == 0
    ev = sorted([(b.start, 1) for b in pb] + [(b.stop, -1) for b in pb])
    alive = peak = 0
    for _, d in ev:
        alive += d
        peak = max(peak, alive)
    assert peak > 40
    m = _model(P, c_walker.segg_seg_len(), c_walker=c_walker, lanes=256, walker_lanes=walker_lanes)
    _same(m.run(mag, chunk_frames=2048), want)


def test_generic_walker_chunk_boundary_inside_a_crowded_stretch(port, synth, c_walker):
    """112 bursts alive where one chunk ends and the next begins (the input of tests/gpu_crowded_boundary_cases.py):
    the list a chunk leaves is the list the next starts from, whatever its length (up to the device's 256)"""
    dense = synth.make_dense_recording(1234)
    rng = np.random.default_rng(77)
    lead = (rng.standard_normal(2_500_000) + 1j * rng.standard_normal(2_500_000)).astype(np.complex64) * np.float32(0.01)
    P, mag, want, nsq = _oracle(port, np.concatenate([lead, dense.iq]))
    assert nsq == 0 and len(want) > 600
    for chunk_frames in (512, 256):
        m = _model(P, c_walker.segg_seg_len(), c_walker=c_walker, lanes=256, walker_lanes=4)
        _same(m.run(mag, chunk_frames=chunk_frames), want)


def test_generic_walker_gives_up_where_it_must(port, synth, c_walker):
    for reason, iq in (("squelch", synth.make_tone_recording(5, 236, 0.02, 0.5)),
                       ("too long", synth.make_tone_recording(6, 1, 0.13, 0.45, total_s=0.75))):
        P, mag, _, _ = _oracle(port, iq)
        with pytest.raises(msm.Bail) as e:
            _model(P, c_walker.segg_seg_len(), c_walker=c_walker, lanes=512).run(mag)
        assert e.value.reason == reason, e.value.reason
