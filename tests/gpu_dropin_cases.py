"""GPU case of the drop-in claim (run in a child process by tests/test_zz_gpu_classify.py): the reference's
unmodified main.c linked against libiridium_b200.so (oracle/_ref/iridium-sniffer-b200) against the reference
program itself (oracle/_ref/iridium-sniffer) on recordings with planted IRA / IBC / IDA frames, in RAW and in
--parsed mode; and the reference's unmodified DETECTOR built with -DUSE_GPU on top of this library's
gpu_burst_fft_* plug-in (oracle/_ref/iridium-sniffer-usegpu) against the same program with --no-gpu.  The reference's own comparison method (test-configurations.sh:150: blank the clock-dependent
fields, sort) on everything that is exact -- ids, payload counts, bit strings, the whole IDA text from the LCW
header on -- and the stated tolerances on the float fields (frequency 2 Hz, magnitude / noise 0.05 dB, confidence
1 %, level 2e-4)."""
import importlib
import importlib.util
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "iridium-sniffer")
NEW_BIN = os.path.join(ROOT, "oracle", "_ref", "iridium-sniffer-b200")
USEGPU_BIN = os.path.join(ROOT, "oracle", "_ref", "iridium-sniffer-usegpu")


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(HERE, name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


fg = _load("frame_gen")


def _run(binary, path, rec, extra):
    r = subprocess.run([binary, "-f", path, "--format=cf32", "-r", str(rec.sample_rate), "-c", str(int(rec.center_freq)),
                        "--file-info=T"] + extra, capture_output=True, text=True, timeout=90)   # (seconds of work; well inside the runner's limit,
                                                                                  # so that a hang is killed HERE, with the program)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.startswith(("RAW:", "IDA:"))]


def _raw_key(l):        # RAW: T ts freq N:mag+noise I:id conf% level npay bits
    f = l.split()
    return (f[5], f[8], f[9] if len(f) > 9 else "")


def _raw_floats(l):
    f = l.split()
    mn = f[4][2:]
    k = max(mn.rfind("+"), mn.rfind("-"))
    return int(f[3]), float(mn[:k]), float(mn[k:]), int(f[6].rstrip("%")), float(f[7])


def _ida_key(l):        # IDA: p-x ts freq conf% level|noise|snr syms dir LCW(...)...
    return (l.split()[6], l.split()[7], l[l.index("LCW("):])


def _ida_floats(l):
    f = l.split()
    lv, no, sn = (float(x) for x in f[5].split("|"))
    return int(f[3]), lv, no, sn, int(f[4].rstrip("%"))


def _compare(ref_lines, new_lines):
    for kind, key, floats, tol in (("RAW:", _raw_key, _raw_floats, (2, 0.05, 0.05, 1, 2e-4)),
                                   ("IDA:", _ida_key, _ida_floats, (2, 0.011, 0.05, 0.05, 1))):
        a = sorted((key(l), floats(l)) for l in ref_lines if l.startswith(kind))
        b = sorted((key(l), floats(l)) for l in new_lines if l.startswith(kind))
        assert [x[0] for x in a] == [x[0] for x in b], kind
        for (ka, fa), (kb, fb) in zip(a, b):
            for va, vb, t in zip(fa, fb, tol):
                assert abs(va - vb) <= t, (kind, ka, fa, fb)
    return len(ref_lines)


def test_reference_main_linked_against_the_library(synth, tmp_path):
    if not (os.path.exists(REF_BIN) and os.path.exists(NEW_BIN)):
        pytest.skip("oracle/_ref programs did not travel with the snapshot")
    n = 0
    for i, (rec, _) in enumerate(fg.planted_recordings(synth)):
        path = str(tmp_path / ("rec%d.cf32" % i))
        rec.iq.tofile(path)
        for extra in ([], ["--parsed"]):
            ref_lines, new_lines = _run(REF_BIN, path, rec, extra), _run(NEW_BIN, path, rec, extra)
            n += _compare(ref_lines, new_lines)
            if extra and i == 1:
                assert sum(l.startswith("IDA:") for l in new_lines) >= 6
    assert n >= 24


def test_reference_detector_on_the_fft_plugin(synth, tmp_path):
    """the reference's own detector, compiled unmodified with -DUSE_GPU, sends its 16-frame batches
    through gpu_burst_fft_create / process / destroy of libiridium_b200.so (burst_detect.c:309,659) and runs its CPU
    state machine on what comes back; the same program with --no-gpu computes the frames itself.  Both FFTs have the
    same arithmetic (the radix-2 DIF stand-in for FFTW is the oracle's, which the CUDA FFT matches bit for bit), so
    the two outputs agree line for line under the reference's own comparison (test-configurations.sh:150)."""
    if not os.path.exists(USEGPU_BIN):
        pytest.skip("oracle/_ref/iridium-sniffer-usegpu did not travel with the snapshot")
    rec = synth.make_recording(1234, duration_s=1.5, n_bursts=12)
    path = str(tmp_path / "rec.cf32")
    rec.iq.tofile(path)

    def run(extra):
        r = subprocess.run([USEGPU_BIN, "-f", path, "--format=cf32", "-r", str(rec.sample_rate), "--file-info=T", "-v"] + extra,
                           capture_output=True, text=True, timeout=120, env=dict(os.environ, IR_PLUGIN_LOG="1"))
        assert r.returncode == 0, r.stderr[-2000:]
        lines = sorted(" ".join(l.split()[3:]) for l in r.stdout.splitlines() if l.startswith("RAW:"))
        return lines, r.stderr

    gpu_lines, gpu_err = run([])
    cpu_lines, cpu_err = run(["--no-gpu"])
    assert "gpu_burst_fft_create(8192, 16)" in gpu_err and "falling back" not in gpu_err, gpu_err[-1500:]   # the plug-in ran
    assert "gpu_burst_fft_create" not in cpu_err
    assert len(cpu_lines) >= 12 and gpu_lines == cpu_lines
