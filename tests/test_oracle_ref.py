"""Pin the CPU restatement (oracle/ir_oracle.c) to the reference.

Three rings, from strongest to most portable:
 1. oracle/_ref/libref_path_dif.so = the reference's own TUs with their FFT calls routed
    through the restatement's radix-2 DIF  -> restatement must match BIT FOR BIT
    (proves every non-FFT operation, incl. FMA placement of the AVX2 build).
 2. oracle/_ref/libref_path.so (independent Stockham FFT shim) -> same bursts, same bits,
    floats within the tolerances of SURVEY.md section 8c.
 3. tests/golden/*.json|npz, produced by the reference binary in the build container
    -> checked everywhere (also on the GPU box, where /root/reference is absent).
Plus the reference's in-tree known answers: unique words / access codes / derived constants.
"""
import json
import os

import numpy as np
import pytest

from oracle import bindings as ob

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _hdr(b):
    return ob.BurstHdr(b["id"], b["start"], b["center_bin"], b["fft_size"], b["sample_rate"],
                       b["magnitude"], b["noise"], b["center_frequency"], b["start_time_ns"])


def _biteq(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.shape == b.shape and a.tobytes() == b.tobytes()


# ----------------------------------------------------------------- known answers
def test_derived_constants_match_reference_verbose_output(port):
    """burst_detect.c:228-235 / burst_downmix.c:243-248 -v output."""
    p = port.det_params()
    assert (p.fft_size, p.burst_width_bins, p.max_bursts) == (8192, 32, 200)
    assert (p.burst_pre_len, p.burst_post_len, p.max_burst_len) == (16384, 160000, 900000)
    assert abs(p.threshold_lin - 4.520657e-02) < 1e-8
    assert p.ringbuf_size == 20_000_000
    q = port.det_params(sample_rate=12_000_000)
    assert (q.fft_size, q.burst_width_bins, q.max_bursts) == (16384, 54, 240)
    assert (q.burst_pre_len, q.burst_post_len, q.max_burst_len) == (32768, 192000, 1080000)
    assert len(port.taps(0)) == 801 and len(port.taps(1)) == 25
    assert len(port.taps(2)) == 20 and len(port.taps(3)) == 51 and len(port.taps(4)) == 51
    _, sl = port.sync_fft(False)
    assert sl == 271


def test_fft_matches_numpy(port):
    rng = np.random.default_rng(0)
    for n in (2, 4, 8, 64, 2048, 4096, 8192, 16384):
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        ref_f = np.fft.fft(x.astype(np.complex128))
        ref_b = np.fft.ifft(x.astype(np.complex128)) * n
        assert np.abs(port.fft(x) - ref_f).max() <= 4e-7 * np.abs(ref_f).max() * np.log2(max(n, 2))
        assert np.abs(port.fft(x, True) - ref_b).max() <= 4e-7 * np.abs(ref_b).max() * np.log2(max(n, 2))


def test_access_code_from_unique_word(port, synth):
    """UW (iridium.h:30-31) differentially decoded == access code (frame_decode.c:51-56)."""
    assert synth.expected_bits(synth.UW_DL) == synth.ACCESS_DL
    assert synth.expected_bits(synth.UW_UL) == synth.ACCESS_UL


# -------------------------------------------------- ring 1: bit-exact with shared FFT
def test_port_bit_exact_with_shared_fft(port, ref_dif, rec_small):
    iq = rec_small.iq
    rb = ref_dif.detect(iq)
    P = port.det_params()
    pb, _, nsq = port.detect(P, iq)
    assert nsq == 0
    assert len(rb) == len(pb) == 13
    n_frames = 0
    for a, b in zip(rb, pb):
        assert (a["id"], a["start"], a["stop"], a["last_active"], a["center_bin"]) == \
               (b.id, b.start, b.stop, b.last_active, b.center_bin)
        assert a["magnitude"] == b.magnitude and a["noise"] == b.noise
        assert _biteq(port.extract(P, iq, b), a["samples"])      # incl. the stale-tail quirk
        rf = ref_dif.downmix(a)
        ok, info, frame, _ = port.downmix(_hdr(a), a["samples"])
        assert ok == (rf is not None)
        if not ok:
            continue
        n_frames += 1
        assert _biteq(frame, rf["samples"])
        assert info.timestamp == rf["timestamp"] and info.center_frequency == rf["center_frequency"]
        assert info.direction == rf["direction"] and info.uw_start == rf["uw_start"]
        for g in (True, False):
            rd = ref_dif.demod(rf, gardner=g)
            ok2, di, bits, llr, _ = port.demod(frame, info.samples_per_symbol,
                                               info.center_frequency, info.direction, gardner=g)
            assert ok2 == (rd is not None)
            if ok2:
                assert _biteq(bits, rd["bits"]) and _biteq(llr, rd["llr"])
                assert di.level == rd["level"] and di.confidence == rd["confidence"]
                assert di.center_frequency == rd["center_frequency"]
                assert di.direction == rd["direction"]
    assert n_frames == 13


@pytest.mark.parametrize("case", ["test_port_detector_equals_reference_on_the_stress_inputs",
                                  "test_port_stages_equal_reference_at_12mhz_on_crowded_traffic"])
def test_port_equals_reference_on_stress_inputs(case):
    """tests/oracle_ref_stress_cases.py, one case per child process: the reference hands out burst extracts whose tails
    are uninitialised malloc memory (SURVEY D10 ii), so what a case leaves on the heap can change what the reference
    does in the next test of the same process (the port is given the same bytes and follows; the tolerance tests that
    use an independent FFT may not)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider",
                        os.path.join("tests", "oracle_ref_stress_cases.py") + "::" + case], cwd=root,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "1 passed" in r.stdout, (r.stdout + r.stderr)[-3000:]


def test_port_kernel_tails_match_reference_avx2(port, ref_dif):
    """Odd lengths exercise the compiler-vectorised remainder loops of simd_avx2.c."""
    import ctypes as C
    L = ref_dif.L
    rng = np.random.default_rng(5)
    b = dict(id=0, start=100000, center_bin=5000, fft_size=8192, sample_rate=10_000_000,
             magnitude=20.0, noise=-100.0, center_frequency=1.622e9, start_time_ns=0)
    for n in (40_000 + 801, 40_000 + 841, 40_000 + 881, 40_000 + 921, 50_123):
        s = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64) * 0.01
        k = np.arange(3000)
        s[9000:12000] += (0.2 * np.exp(2j * np.pi * 0.11 * k)).astype(np.complex64)
        b["samples"] = s
        rf = ref_dif.downmix(b)
        ok, info, frame, _ = port.downmix(_hdr(b), s)
        assert ok == (rf is not None)
        if ok:
            assert _biteq(frame, rf["samples"])


# ------------------------------------------- ring 2: independent FFT, tolerances of 8c
def test_port_vs_reference_independent_fft(port, ref, rec_small):
    iq = rec_small.iq
    rb = ref.detect(iq)
    P = port.det_params()
    pb, _, _ = port.detect(P, iq)
    assert [(a["id"], a["start"], a["stop"], a["center_bin"]) for a in rb] == \
           [(b.id, b.start, b.stop, b.center_bin) for b in pb]
    truth = {t.bits for t in rec_small.truth}
    n_ok = 0
    for a, b in zip(rb, pb):
        assert abs(a["magnitude"] - b.magnitude) < 0.05 and abs(a["noise"] - b.noise) < 0.05
        rf = ref.downmix(a)
        ok, info, frame, _ = port.downmix(_hdr(a), a["samples"])
        assert ok == (rf is not None)
        assert info.timestamp == rf["timestamp"]
        assert abs(info.center_frequency - rf["center_frequency"]) < 2.0
        rd = ref.demod(rf)
        ok2, di, bits, _, _ = port.demod(frame, 10.0, info.center_frequency, info.direction)
        assert ok2 == (rd is not None)
        if ok2:
            n_ok += 1
            assert _biteq(bits, rd["bits"])
            assert "".join(map(str, bits)) in truth
            assert abs(di.level - rd["level"]) < 2e-4 and abs(di.confidence - rd["confidence"]) <= 1
            assert abs(di.center_frequency - rd["center_frequency"]) < 2.0
    assert n_ok == 12


# ------------------------------------------------------------- ring 3: golden vectors
def _match_lines(results, gold_lines):
    """Match by burst id (detection sets agree on these fixtures)."""
    by_id = {r["id"]: r for r in results}
    for g in gold_lines:
        assert g["id"] in by_id, f"burst {g['id']} missing"
        r = by_id[g["id"]]
        assert "".join(map(str, r["bits"])) == g["bits"]
        assert r["n_payload_symbols"] == g["n_payload"]
        assert abs(int(r["center_frequency"] + 0.5) - g["freq_hz"]) <= 2
        assert abs(r["magnitude"] - g["magnitude"]) <= 0.05 and abs(r["noise"] - g["noise"]) <= 0.05
        assert abs(r["confidence"] - g["confidence"]) <= 1
        assert abs(r["level"] - g["level"]) <= 2e-4
    assert len(results) == len(gold_lines)


def test_port_against_golden_reference_lines(port, rec_small):
    gold = json.load(open(os.path.join(GOLD, "ref_lines.json")))
    res, stats = port.run(rec_small.iq)
    assert stats["n_bursts"] == 13
    _match_lines(res, gold["config1_cf32_10MHz_seed1234"]["lines"])
    res_ng, _ = port.run(rec_small.iq, gardner=False)
    _match_lines(res_ng, gold["config1_no_gardner"]["lines"])


def test_port_against_golden_ci16_12mhz(port, synth):
    gold = json.load(open(os.path.join(GOLD, "ref_lines.json")))["config3_ci16_12MHz_seed3"]
    g = gold["gen"]
    rec = synth.make_recording(g["seed"], sample_rate=g["sample_rate"], duration_s=g["duration_s"],
                               n_bursts=g["n_bursts"], fmt="ci16", center_freq=g["center_freq"])
    iq = port.convert_ci16(rec.iq)
    res, stats = port.run(iq, center_frequency=g["center_freq"], sample_rate=g["sample_rate"])
    assert stats["n_bursts"] == 9
    _match_lines(res, gold["lines"])


def test_port_against_golden_stage_vectors(port, rec_small):
    v = np.load(os.path.join(GOLD, "config1_stage_vectors.npz"))
    P = port.det_params()
    pb, _, _ = port.detect(P, rec_small.iq)
    for b in pb[:2]:
        hdr = v[f"burst{b.id}_hdr"]
        assert (b.id, b.start, b.stop, b.center_bin) == tuple(int(x) for x in hdr)
        s = port.extract(P, rec_small.iq, b)
        h = ob.BurstHdr(b.id, b.start, b.center_bin, P.fft_size, P.sample_rate, b.magnitude,
                        b.noise, P.center_frequency, 0)
        ok, info, frame, _ = port.downmix(h, s)
        gf = v[f"burst{b.id}_frame"]
        assert ok and frame.shape == gf.shape
        assert np.abs(frame - gf).max() <= 1e-4 * np.abs(gf).max()
        ok2, di, bits, _, _ = port.demod(frame, 10.0, info.center_frequency, info.direction)
        assert ok2 and _biteq(bits, v[f"burst{b.id}_bits"])
        sc = v[f"burst{b.id}_scalars"]
        assert abs(info.center_frequency - sc[0]) < 2 and abs(di.center_frequency - sc[1]) < 2
        assert abs(di.level - sc[2]) < 2e-4 and abs(b.magnitude - sc[4]) < 0.05


def test_raw_line_format(port):
    """frame_output.c:182-192 format string, checked on a golden line."""
    gold = json.load(open(os.path.join(GOLD, "ref_lines.json")))["config1_cf32_10MHz_seed1234"]["lines"][0]
    r = dict(timestamp=994_187_000 + 5_000_000_000, center_frequency=gold["freq_hz"] + 0.2,
             magnitude=gold["magnitude"], noise=gold["noise"], id=gold["id"],
             confidence=gold["confidence"], level=gold["level"],
             n_payload_symbols=gold["n_payload"],
             bits=np.array([int(c) for c in gold["bits"]], np.uint8))
    line = port.format_raw("T", 5_000_000_000, r)
    f = line.split()
    assert f[0] == "RAW:" and f[1] == "T" and f[2] == "0000994.1870"
    assert f[3] == "%010d" % gold["freq_hz"]
    assert f[4] == "N:%05.2f%+06.2f" % (gold["magnitude"], gold["noise"])
    assert f[5] == "I:%011d" % gold["id"] and f[6] == "100%" and f[8] == "179"
    assert f[9] == gold["bits"] and line.endswith("\n")
