"""The oracle port against the reference's own stage functions (oracle/_ref, compiled unmodified) on the inputs the
GPU's state machines are stressed with.  Not collected by name: tests/test_oracle_ref.py runs each case in a child
process (see there why)."""
import os

import numpy as np

from test_oracle_ref import _biteq, _hdr


def test_port_detector_equals_reference_on_the_stress_inputs(port, ref_dif, synth):
    """The inputs the GPU's state machines are stressed with (tests/test_gpu_detector_stress.py holds the device against
    the PORT there): the reference's own detector, compiled unmodified, gives the port's burst list field for field on
    dense traffic (674 bursts, ~170 alive at once), squelch, a carrier longer than max_burst_len (forced baseline
    update), and 12 MHz / 16384-pt clipped int8 samples (229 bursts from 14 planted, 76 alive at once)."""
    def same(iq, fs=10_000_000, fmt="cf32", iq_port=None, n=None):
        rb = ref_dif.detect(iq, sample_rate=fs, fmt=fmt)
        pb, _, _ = port.detect(port.det_params(sample_rate=fs), iq if iq_port is None else iq_port)
        a = [(x["id"], x["start"], x["stop"], x["last_active"], x["center_bin"], x["magnitude"], x["noise"]) for x in rb]
        b = [(x.id, x.start, x.stop, x.last_active, x.center_bin, x.magnitude, x.noise) for x in pb]
        assert a == b
        if n is not None:
            assert len(a) == n
    same(synth.make_dense_recording(1234).iq, n=674)
    same(synth.make_tone_recording(5, 236, 0.02, 0.5), n=0)                     # 236 carriers > max_bursts: squelched
    same(synth.make_tone_recording(6, 1, 0.13, 0.45, total_s=0.75), n=4)
    rec = synth.make_recording(3, sample_rate=12_000_000, duration_s=1.6, n_bursts=14, fmt="ci16", snr_db=(26.0, 30.0))
    same((rec.iq >> 8).astype(np.int8), fs=12_000_000, fmt="ci8", iq_port=port.convert_ci16(rec.iq), n=229)


def test_port_stages_equal_reference_at_12mhz_on_crowded_traffic(port, ref_dif, synth):
    """12 MHz (decimation 48) and 229 bursts, most of them spurious detections around clipped strong bursts: every
    burst through the reference's own burst_downmix and qpsk_demod and through the port -- same verdict, same
    250 kHz frame, same bits / LLRs / level / confidence (shared FFT: bit for bit)."""
    from oracle import bindings as ob
    # (a burst_downmix object takes its decimation from the first burst it sees, the reference's and the port's alike:
    #  fresh ones for 12 MHz, not the session's that other tests have used at 10 MHz)
    ref_dif, port = ob.Ref(ob.REF_DIF_SO), ob.Port()
    rec = synth.make_recording(3, sample_rate=12_000_000, duration_s=1.6, n_bursts=14, fmt="ci16", snr_db=(26.0, 30.0))
    rb = ref_dif.detect((rec.iq >> 8).astype(np.int8), sample_rate=12_000_000, fmt="ci8")
    assert len(rb) == 229
    n_dm = n_fr = 0
    for a in rb:
        rf = ref_dif.downmix(a)
        ok, info, frame, _ = port.downmix(_hdr(a), a["samples"])
        assert ok == (rf is not None)
        if not ok:
            continue
        n_dm += 1
        assert _biteq(frame, rf["samples"]) and info.timestamp == rf["timestamp"] and info.direction == rf["direction"]
        assert info.center_frequency == rf["center_frequency"] and info.uw_start == rf["uw_start"]
        rd = ref_dif.demod(rf)
        ok2, di, bits, llr, _ = port.demod(frame, info.samples_per_symbol, info.center_frequency, info.direction)
        assert ok2 == (rd is not None)
        if ok2:
            n_fr += 1
            assert _biteq(bits, rd["bits"]) and _biteq(llr, rd["llr"])
            assert di.level == rd["level"] and di.confidence == rd["confidence"]
            assert di.center_frequency == rd["center_frequency"] and di.direction == rd["direction"]
    # (how many survive depends on what the reference's malloc hands its ring buffer: the extracts' tails past the
    #  last fed sample are uninitialised memory, SURVEY D10 (ii) -- both sides are given the same bytes)
    print("downmix frames", n_dm, "demod frames", n_fr)
    assert n_dm >= 50 and n_fr >= 10
