"""The reference-shaped per-item API and the plug-in ABI (include/ir_ref_api.h, include/burst_fft.h)
driven exactly the way main.c's threads drive the reference: 32768-sample feed calls, one burst
at a time through burst_downmix_process, one frame at a time through qpsk_demod."""
import ctypes as C
import importlib

import numpy as np
import pytest

from oracle import bindings as ob

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    pl = importlib.import_module("iridium-sniffer_b200.pipeline")
    L = pl.load_library()
    L.burst_detector_create.restype = C.c_void_p
    L.burst_detector_create.argtypes = [C.POINTER(ob.RefBurstConfig)]
    L.burst_detector_destroy.argtypes = [C.c_void_p]
    L.burst_detector_total_count.restype = C.c_uint64
    L.burst_detector_total_count.argtypes = [C.c_void_p]
    L.burst_detector_noise_floor.restype = C.c_float
    L.burst_detector_noise_floor.argtypes = [C.c_void_p]
    L.burst_downmix_create.restype = C.c_void_p
    L.burst_downmix_create.argtypes = [C.c_void_p]
    L.burst_downmix_destroy.argtypes = [C.c_void_p]
    L.burst_downmix_process.restype = C.c_int
    L.burst_downmix_process.argtypes = [C.c_void_p, C.POINTER(ob.RefBurstData),
                                        C.POINTER(C.POINTER(ob.RefDownmixFrame))]
    L.qpsk_demod.restype = C.c_int
    L.qpsk_demod.argtypes = [C.POINTER(ob.RefDownmixFrame), C.POINTER(C.POINTER(ob.RefDemodFrame))]
    L.gpu_burst_fft_create.restype = C.c_void_p
    L.gpu_burst_fft_create.argtypes = [C.c_int, C.c_int, C.c_void_p]
    L.gpu_burst_fft_process.restype = C.c_int
    L.gpu_burst_fft_process.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.gpu_burst_fft_destroy.argtypes = [C.c_void_p]
    return L


CB = C.CFUNCTYPE(None, C.POINTER(ob.RefBurstData), C.c_void_p)
libc = C.CDLL(None)
libc.free.argtypes = [C.c_void_p]


def cfree(p):
    libc.free(C.cast(p, C.c_void_p))


def test_plugin_fft_matches_oracle(lib, port, rec_small):
    N, B = 8192, 16
    win = port.det_window(N)
    g = lib.gpu_burst_fft_create(N, B, win.ctypes.data_as(C.c_void_p))
    assert g
    x = np.ascontiguousarray(rec_small.iq[512 * N:(512 + B) * N])
    out = np.empty((B, N), np.float32)
    assert lib.gpu_burst_fft_process(g, x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), B) == 0
    for k in range(B):
        assert out[k].tobytes() == port.frame_mag(x[k * N:(k + 1) * N], win).tobytes()
    assert lib.gpu_burst_fft_process(g, x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), B + 1) == -1
    lib.gpu_burst_fft_destroy(g)


def test_reference_shaped_stages(lib, port, rec_small):
    iq = rec_small.iq
    cfg = ob.RefBurstConfig(1_622_000_000.0, 10_000_000, 0, 0, 0, 40000, 0, 0, 16.0, 512, 1)
    det = lib.burst_detector_create(C.byref(cfg))
    assert det
    got = []

    @CB
    def cb(bp, user):
        b = bp.contents
        s = np.ctypeslib.as_array(C.cast(b.samples, C.POINTER(C.c_float)), (2 * b.num_samples,)).copy().view(np.complex64)
        got.append(dict(id=b.info.id, start=b.info.start, stop=b.info.stop, last_active=b.info.last_active,
                        center_bin=b.info.center_bin, magnitude=b.info.magnitude, noise=b.info.noise,
                        center_frequency=b.center_frequency, sample_rate=b.sample_rate, fft_size=b.fft_size,
                        start_time_ns=b.start_time_ns, samples=s))
        cfree(b.samples)              # ownership passed to the callback (burst_detect.h:71-72)
        cfree(bp)

    flat = iq.view(np.float32)
    for off in range(0, iq.shape[0], 32768):      # main.c:225
        m = min(32768, iq.shape[0] - off)
        lib.burst_detector_feed_cf32(det, flat[2 * off:].ctypes.data_as(C.c_void_p), m, cb, None)
    assert lib.burst_detector_total_count(det) == 13
    assert -120.0 < lib.burst_detector_noise_floor(det) < 0.0
    lib.burst_detector_destroy(det)

    P = port.det_params()
    pb, _, _ = port.detect(P, iq)
    assert len(got) == len(pb)
    dm = lib.burst_downmix_create(None)
    assert dm
    n_ok = 0
    for g, o in zip(got, pb):
        assert (g["id"], g["start"], g["stop"], g["center_bin"]) == (o.id, o.start, o.stop, o.center_bin)
        assert g["magnitude"] == o.magnitude and g["noise"] == o.noise
        assert g["samples"].tobytes() == port.extract(P, iq, o).tobytes()
        s = np.ascontiguousarray(g["samples"])
        bd = ob.RefBurstData()
        bd.info = ob.RefBurstInfo(g["id"], g["start"], g["stop"], g["last_active"], g["center_bin"],
                                  g["magnitude"], g["noise"])
        bd.center_frequency, bd.sample_rate, bd.fft_size = g["center_frequency"], g["sample_rate"], g["fft_size"]
        bd.start_time_ns, bd.num_samples, bd.samples = g["start_time_ns"], s.shape[0], s.ctypes.data
        fr = C.POINTER(ob.RefDownmixFrame)()
        n = lib.burst_downmix_process(dm, C.byref(bd), C.byref(fr))
        hdr = ob.BurstHdr(g["id"], g["start"], g["center_bin"], g["fft_size"], g["sample_rate"], g["magnitude"],
                          g["noise"], g["center_frequency"], g["start_time_ns"])
        ok, info, frame, _ = port.downmix(hdr, s)
        assert (n == 1) == ok
        if not ok:
            continue
        f = fr.contents
        fs = np.ctypeslib.as_array(C.cast(f.samples, C.POINTER(C.c_float)), (2 * f.num_samples,)).copy().view(np.complex64)
        assert f.timestamp == info.timestamp and f.direction == info.direction and f.num_samples == info.num_samples
        assert abs(f.center_frequency - info.center_frequency) < 1e-3
        assert np.abs(fs - frame).max() <= 2e-5 * np.abs(frame).max()
        out = C.POINTER(ob.RefDemodFrame)()
        okd = lib.qpsk_demod(fr, C.byref(out))
        ok2, di, bits, llr, _ = port.demod(frame, 10.0, info.center_frequency, info.direction)
        assert bool(okd) == ok2
        if ok2:
            n_ok += 1
            d = out.contents
            assert np.ctypeslib.as_array(d.bits, (d.n_bits,)).tobytes() == bits.tobytes()
            assert d.n_symbols == di.n_symbols and d.confidence == di.confidence
            assert abs(d.level - di.level) < 2e-6 and abs(d.center_frequency - di.center_frequency) < 0.05
            cfree(d.bits); cfree(d.llr); cfree(out)
        cfree(f.samples); cfree(fr)
    assert n_ok == 12
    lib.burst_downmix_destroy(dm)
