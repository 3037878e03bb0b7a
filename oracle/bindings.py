"""ctypes bindings for the CPU oracles -- TEST INFRASTRUCTURE ONLY.

Two libraries live under oracle/:
  * libir_oracle.so        our C restatement (oracle/ir_oracle.c)            -> class Port
  * _ref/libref_path.so    the reference's own TUs, unmodified, + FFT shim    -> class Ref
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module.  The product never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libir_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref_path.so")
REF_DBL_SO = os.path.join(HERE, "_ref", "libref_path_dbl.so")
REF_DIF_SO = os.path.join(HERE, "_ref", "libref_path_dif.so")
REF_BIN = os.path.join(HERE, "_ref", "iridium-sniffer")


def build(port: bool = True, ref: bool = True) -> None:
    """Compile the oracle libraries (ref only when /root/reference is present)."""
    targets = (["port"] if port else []) + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


# --------------------------------------------------------------------- port structs
class DetParams(C.Structure):
    _fields_ = [("center_frequency", C.c_double), ("sample_rate", C.c_int), ("fft_size", C.c_int),
                ("burst_pre_len", C.c_int), ("burst_post_len", C.c_int),
                ("burst_width_bins", C.c_int), ("max_bursts", C.c_int),
                ("max_burst_len", C.c_int), ("history_size", C.c_int),
                ("threshold_db", C.c_float), ("threshold_lin", C.c_float),
                ("ringbuf_size", C.c_size_t)]


class OrcBurst(C.Structure):
    _fields_ = [("id", C.c_uint64), ("start", C.c_uint64), ("stop", C.c_uint64),
                ("last_active", C.c_uint64), ("center_bin", C.c_int32),
                ("magnitude", C.c_float), ("noise", C.c_float), ("peak_rel", C.c_float),
                ("base_at_create", C.c_float), ("emit_count", C.c_uint64),
                ("ring_start", C.c_uint64)]


class BurstHdr(C.Structure):
    _fields_ = [("id", C.c_uint64), ("start", C.c_uint64), ("center_bin", C.c_int32),
                ("fft_size", C.c_int32), ("sample_rate", C.c_int32), ("magnitude", C.c_float),
                ("noise", C.c_float), ("center_frequency", C.c_double),
                ("start_time_ns", C.c_uint64)]


class FrameInfo(C.Structure):
    _fields_ = [("ok", C.c_int32), ("fail_stage", C.c_int32), ("id", C.c_uint64),
                ("timestamp", C.c_uint64), ("center_frequency", C.c_double),
                ("sample_rate", C.c_float), ("samples_per_symbol", C.c_float),
                ("direction", C.c_int32), ("magnitude", C.c_float), ("noise", C.c_float),
                ("uw_start", C.c_float), ("num_samples", C.c_int32), ("dec_len", C.c_int32),
                ("start", C.c_int32), ("center_offset", C.c_float), ("cfo_peak_bin", C.c_int32),
                ("corr_offset", C.c_int32), ("uw_start_idx", C.c_int32),
                ("corr_re", C.c_float), ("corr_im", C.c_float),
                ("incr_coarse_re", C.c_float), ("incr_coarse_im", C.c_float),
                ("incr_fine_re", C.c_float), ("incr_fine_im", C.c_float)]


class DemodInfo(C.Structure):
    _fields_ = [("ok", C.c_int32), ("direction", C.c_int32), ("confidence", C.c_int32),
                ("level", C.c_float), ("n_symbols", C.c_int32), ("n_payload_symbols", C.c_int32),
                ("n_bits", C.c_int32), ("center_frequency", C.c_double),
                ("total_phase", C.c_float), ("n_raw_symbols", C.c_int32)]


class OrcResult(C.Structure):
    _fields_ = [("id", C.c_uint64), ("timestamp", C.c_uint64), ("center_frequency", C.c_double),
                ("direction", C.c_int32), ("magnitude", C.c_float), ("noise", C.c_float),
                ("confidence", C.c_int32), ("level", C.c_float), ("n_symbols", C.c_int32),
                ("n_payload_symbols", C.c_int32), ("n_bits", C.c_int32),
                ("bits_offset", C.c_uint32)]


class OrcRun(C.Structure):
    _fields_ = [("n_bursts", C.c_size_t), ("n_frames", C.c_size_t), ("n_results", C.c_size_t),
                ("results", C.POINTER(OrcResult)), ("bits", C.POINTER(C.c_uint8)),
                ("bits_len", C.c_size_t), ("t_detect_s", C.c_double),
                ("t_downmix_s", C.c_double), ("t_demod_s", C.c_double)]


def _cf(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.complex64)
    return a


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Port:
    """oracle/libir_oracle.so."""

    def __init__(self, path: str = PORT_SO):
        if not os.path.exists(path):
            build(port=True, ref=False)
        L = self.L = C.CDLL(path)
        L.orc_detect.restype = C.c_size_t
        L.orc_detect.argtypes = [C.POINTER(DetParams), C.c_void_p, C.c_size_t, C.c_size_t,
                                 C.POINTER(C.POINTER(OrcBurst)), C.c_void_p, C.POINTER(C.c_int)]
        L.orc_burst_num_samples.restype = C.c_size_t
        L.orc_burst_num_samples.argtypes = [C.POINTER(DetParams), C.POINTER(OrcBurst)]
        L.orc_burst_extract.restype = C.c_size_t
        L.orc_burst_extract.argtypes = [C.POINTER(DetParams), C.c_void_p, C.c_size_t,
                                        C.POINTER(OrcBurst), C.c_void_p]
        L.orc_det_params_init.argtypes = [C.POINTER(DetParams), C.c_double, C.c_int, C.c_int,
                                          C.c_int, C.c_float]
        L.orc_downmix_create.restype = C.c_void_p
        L.orc_downmix_destroy.argtypes = [C.c_void_p]
        L.orc_downmix_taps.restype = C.POINTER(C.c_float)
        L.orc_downmix_taps.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.orc_downmix_sync_fft.restype = C.c_void_p
        L.orc_downmix_sync_fft.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.orc_downmix_cfo_window.restype = C.POINTER(C.c_float)
        L.orc_downmix_cfo_window.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.orc_downmix_process.restype = C.c_int
        L.orc_downmix_process.argtypes = [C.c_void_p, C.POINTER(BurstHdr), C.c_void_p, C.c_size_t,
                                          C.POINTER(FrameInfo), C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p]
        L.orc_demod.restype = C.c_int
        L.orc_demod.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_int, C.c_int,
                                C.POINTER(DemodInfo), C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_fft.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_det_window.argtypes = [C.c_void_p, C.c_int]
        L.orc_det_frame_mag.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_convert_ci8.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_convert_ci16.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_format_raw.restype = C.c_int
        L.orc_format_raw.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_uint64, C.c_uint64,
                                     C.c_double, C.c_float, C.c_float, C.c_uint64, C.c_int,
                                     C.c_float, C.c_int, C.c_void_p, C.c_int]
        L.orc_run_recording.restype = C.c_int
        L.orc_run_recording.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.c_int, C.c_float,
                                        C.c_size_t, C.c_uint64, C.c_int, C.POINTER(OrcRun)]
        L.orc_run_free.argtypes = [C.POINTER(OrcRun)]
        L.orc_free.argtypes = [C.c_void_p]
        self._dm = None

    # ---- primitives
    def fft(self, x: np.ndarray, inverse: bool = False) -> np.ndarray:
        y = _cf(x).copy()
        self.L.orc_fft(_p(y), y.shape[0], int(inverse))
        return y

    def det_params(self, center_frequency=1_622_000_000.0, sample_rate=10_000_000, fft_size=0,
                   burst_width_hz=0, threshold_db=16.0) -> DetParams:
        p = DetParams()
        self.L.orc_det_params_init(C.byref(p), center_frequency, sample_rate, fft_size,
                                   burst_width_hz, threshold_db)
        return p

    def det_window(self, n: int) -> np.ndarray:
        w = np.empty(n, np.float32)
        self.L.orc_det_window(_p(w), n)
        return w

    def frame_mag(self, frame: np.ndarray, window: np.ndarray) -> np.ndarray:
        f = _cf(frame)
        out = np.empty(f.shape[0], np.float32)
        self.L.orc_det_frame_mag(_p(f), _p(window), f.shape[0], _p(out))
        return out

    def convert_ci8(self, iq: np.ndarray) -> np.ndarray:
        iq = np.ascontiguousarray(iq, np.int8)
        out = np.empty(iq.shape[0] // 2, np.complex64)
        self.L.orc_convert_ci8(_p(iq), out.shape[0], _p(out))
        return out

    def convert_ci16(self, iq: np.ndarray) -> np.ndarray:
        iq = np.ascontiguousarray(iq, np.int16)
        out = np.empty(iq.shape[0] // 2, np.complex64)
        self.L.orc_convert_ci16(_p(iq), out.shape[0], _p(out))
        return out

    # ---- detector
    def detect(self, params: DetParams, iq: np.ndarray, feed_block: int = 32768,
               dump_mag: bool = False):
        iq = _cf(iq)
        out = C.POINTER(OrcBurst)()
        nsq = C.c_int(0)
        mag = None
        if dump_mag:
            nf = iq.shape[0] // params.fft_size
            mag = np.zeros((nf, params.fft_size), np.float32)
        n = self.L.orc_detect(C.byref(params), _p(iq), iq.shape[0], feed_block, C.byref(out),
                              _p(mag) if mag is not None else None, C.byref(nsq))
        bursts = []
        for i in range(n):
            b = OrcBurst()
            C.memmove(C.byref(b), C.byref(out[i]), C.sizeof(OrcBurst))
            bursts.append(b)
        if n:
            self.L.orc_free(out)
        return bursts, mag, nsq.value

    def extract(self, params: DetParams, iq: np.ndarray, b: OrcBurst) -> np.ndarray:
        iq = _cf(iq)
        n = self.L.orc_burst_num_samples(C.byref(params), C.byref(b))
        dst = np.empty(n, np.complex64)
        self.L.orc_burst_extract(C.byref(params), _p(iq), iq.shape[0], C.byref(b), _p(dst))
        return dst

    # ---- downmix
    @property
    def dm(self):
        if self._dm is None:
            self._dm = self.L.orc_downmix_create()
        return self._dm

    def taps(self, which: int) -> np.ndarray:
        n = C.c_int(0)
        p = self.L.orc_downmix_taps(self.dm, which, C.byref(n))
        return np.ctypeslib.as_array(p, (n.value,)).copy()

    def sync_fft(self, uplink: bool):
        n = C.c_int(0)
        p = self.L.orc_downmix_sync_fft(self.dm, int(uplink), C.byref(n))
        arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), (4096,)).copy()
        return arr.view(np.complex64), n.value

    def cfo_window(self) -> np.ndarray:
        n = C.c_int(0)
        p = self.L.orc_downmix_cfo_window(self.dm, C.byref(n))
        return np.ctypeslib.as_array(p, (n.value,)).copy()

    def downmix(self, hdr: BurstHdr, samples: np.ndarray, trace: bool = False):
        s = _cf(samples)
        info = FrameInfo()
        frame = np.zeros(4480, np.complex64)
        cap = max(s.shape[0] // 20 + 16, 16)
        dec = np.zeros(cap, np.complex64) if trace else None
        nl = np.zeros(cap, np.complex64) if trace else None
        rr = np.zeros(cap, np.complex64) if trace else None
        ok = self.L.orc_downmix_process(self.dm, C.byref(hdr), _p(s), s.shape[0], C.byref(info),
                                        _p(frame), _p(dec) if trace else None,
                                        _p(nl) if trace else None, _p(rr) if trace else None)
        tr = None
        if trace:
            tr = {"dec": dec[:info.dec_len], "nlpf": nl[:info.dec_len],
                  "rrc": rr[:max(info.dec_len - info.start, 0)]}
        return bool(ok), info, frame[:info.num_samples].copy(), tr

    # ---- demod
    def demod(self, frame: np.ndarray, sps: float, center_frequency: float, direction: int,
              gardner: bool = True):
        f = _cf(frame)
        info = DemodInfo()
        bits = np.zeros(900, np.uint8)
        llr = np.zeros(900, np.float32)
        pll = np.zeros(f.shape[0] // 10 + 2, np.complex64)
        ok = self.L.orc_demod(_p(f), f.shape[0], sps, center_frequency, direction, int(gardner),
                              C.byref(info), _p(bits), _p(llr), _p(pll))
        return bool(ok), info, bits[:info.n_bits].copy(), llr[:info.n_bits].copy(), \
            pll[:info.n_raw_symbols].copy()

    def format_raw(self, file_info: str, t0: int, r) -> str:
        buf = C.create_string_buffer(2048)
        bits = np.ascontiguousarray(r["bits"], np.uint8)
        self.L.orc_format_raw(buf, 2048, file_info.encode(), t0, r["timestamp"],
                              r["center_frequency"], r["magnitude"], r["noise"], r["id"],
                              r["confidence"], r["level"], r["n_payload_symbols"], _p(bits),
                              bits.shape[0])
        return buf.value.decode()

    # ---- whole path
    def run(self, iq: np.ndarray, center_frequency=1_622_000_000.0, sample_rate=10_000_000,
            threshold_db=16.0, feed_block=32768, start_time_ns=0, gardner=True):
        iq = _cf(iq)
        run = OrcRun()
        self.L.orc_run_recording(_p(iq), iq.shape[0], center_frequency, sample_rate, threshold_db,
                                 feed_block, start_time_ns, int(gardner), C.byref(run))
        res = []
        allbits = np.ctypeslib.as_array(run.bits, (max(run.bits_len, 1),))
        for i in range(run.n_results):
            r = run.results[i]
            res.append(dict(id=r.id, timestamp=r.timestamp, center_frequency=r.center_frequency,
                            direction=r.direction, magnitude=r.magnitude, noise=r.noise,
                            confidence=r.confidence, level=r.level, n_symbols=r.n_symbols,
                            n_payload_symbols=r.n_payload_symbols,
                            bits=allbits[r.bits_offset:r.bits_offset + r.n_bits].copy()))
        stats = dict(n_bursts=run.n_bursts, n_frames=run.n_frames, n_results=run.n_results,
                     t_detect_s=run.t_detect_s, t_downmix_s=run.t_downmix_s,
                     t_demod_s=run.t_demod_s)
        self.L.orc_run_free(C.byref(run))
        return res, stats


# ------------------------------------------------------------ reference structs
class RefBurstInfo(C.Structure):           # burst_detect.h:29-37
    _fields_ = [("id", C.c_uint64), ("start", C.c_uint64), ("stop", C.c_uint64),
                ("last_active", C.c_uint64), ("center_bin", C.c_int), ("magnitude", C.c_float),
                ("noise", C.c_float)]


class RefBurstData(C.Structure):           # burst_detect.h:40-48
    _fields_ = [("info", RefBurstInfo), ("center_frequency", C.c_double),
                ("sample_rate", C.c_int), ("fft_size", C.c_int), ("start_time_ns", C.c_uint64),
                ("num_samples", C.c_size_t), ("samples", C.c_void_p)]


class RefBurstConfig(C.Structure):         # burst_detect.h:51-63
    _fields_ = [("center_frequency", C.c_double), ("sample_rate", C.c_int), ("fft_size", C.c_int),
                ("burst_pre_len", C.c_int), ("burst_post_len", C.c_int), ("burst_width", C.c_int),
                ("max_bursts", C.c_int), ("max_burst_len", C.c_int), ("threshold", C.c_float),
                ("history_size", C.c_int), ("use_gpu", C.c_int)]


class RefDownmixFrame(C.Structure):        # burst_downmix.h:39-51
    _fields_ = [("id", C.c_uint64), ("timestamp", C.c_uint64), ("center_frequency", C.c_double),
                ("sample_rate", C.c_float), ("samples_per_symbol", C.c_float),
                ("direction", C.c_int), ("magnitude", C.c_float), ("noise", C.c_float),
                ("uw_start", C.c_float), ("num_samples", C.c_size_t), ("samples", C.c_void_p)]


class RefDemodFrame(C.Structure):          # qpsk_demod.h:24-38
    _fields_ = [("id", C.c_uint64), ("timestamp", C.c_uint64), ("center_frequency", C.c_double),
                ("direction", C.c_int), ("magnitude", C.c_float), ("noise", C.c_float),
                ("confidence", C.c_int), ("level", C.c_float), ("n_symbols", C.c_int),
                ("n_payload_symbols", C.c_int), ("bits", C.POINTER(C.c_uint8)),
                ("llr", C.POINTER(C.c_float)), ("n_bits", C.c_int)]


class Ref:
    """oracle/_ref/libref_path.so -- the reference's own stage functions."""

    def __init__(self, path: str = REF_SO, no_simd: bool = False, gardner: bool = True):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = self.L = C.CDLL(path)
        L.ref_init.argtypes = [C.c_int, C.c_int, C.c_int]
        L.ref_init(int(no_simd), int(gardner), 0)
        L.burst_detector_create.restype = C.c_void_p
        L.burst_detector_create.argtypes = [C.POINTER(RefBurstConfig)]
        L.burst_detector_destroy.argtypes = [C.c_void_p]
        L.ref_burst_list_new.restype = C.c_void_p
        L.ref_burst_list_len.restype = C.c_size_t
        L.ref_burst_list_len.argtypes = [C.c_void_p]
        L.ref_burst_list_get.restype = C.POINTER(RefBurstData)
        L.ref_burst_list_get.argtypes = [C.c_void_p, C.c_size_t]
        L.ref_burst_list_free.argtypes = [C.c_void_p]
        L.ref_detect_cf32.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
        L.ref_detect_ci8.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
        L.burst_downmix_create.restype = C.c_void_p
        L.burst_downmix_create.argtypes = [C.c_void_p]
        L.burst_downmix_process.restype = C.c_int
        L.burst_downmix_process.argtypes = [C.c_void_p, C.POINTER(RefBurstData),
                                            C.POINTER(C.POINTER(RefDownmixFrame))]
        L.qpsk_demod.restype = C.c_int
        L.qpsk_demod.argtypes = [C.POINTER(RefDownmixFrame), C.POINTER(C.POINTER(RefDemodFrame))]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_set_gardner.argtypes = [C.c_int]
        self._dm = None

    def detect(self, iq: np.ndarray, center_frequency=1_622_000_000.0, sample_rate=10_000_000,
               threshold_db=16.0, feed_block=32768, fmt: str = "cf32"):
        """Returns list of dicts with the burst_data_t fields and a copy of the samples."""
        cfg = RefBurstConfig(center_frequency, sample_rate, 0, 0, 0, 40000, 0, 0, threshold_db,
                             512, 0)
        det = self.L.burst_detector_create(C.byref(cfg))
        lst = self.L.ref_burst_list_new()
        if fmt == "cf32":
            a = _cf(iq)
            self.L.ref_detect_cf32(det, _p(a), a.shape[0], feed_block, lst)
        else:
            a = np.ascontiguousarray(iq, np.int8)
            self.L.ref_detect_ci8(det, _p(a), a.shape[0] // 2, feed_block, lst)
        out = []
        for i in range(self.L.ref_burst_list_len(lst)):
            b = self.L.ref_burst_list_get(lst, i).contents
            s = np.ctypeslib.as_array(C.cast(b.samples, C.POINTER(C.c_float)),
                                      (2 * b.num_samples,)).copy().view(np.complex64)
            out.append(dict(id=b.info.id, start=b.info.start, stop=b.info.stop,
                            last_active=b.info.last_active, center_bin=b.info.center_bin,
                            magnitude=b.info.magnitude, noise=b.info.noise,
                            center_frequency=b.center_frequency, sample_rate=b.sample_rate,
                            fft_size=b.fft_size, start_time_ns=b.start_time_ns, samples=s))
        self.L.ref_burst_list_free(lst)
        # burst_detector_destroy prints to stderr; keep it (it also frees ~180 MB)
        self.L.burst_detector_destroy(det)
        return out

    @property
    def dm(self):
        if self._dm is None:
            self._dm = self.L.burst_downmix_create(None)
        return self._dm

    def downmix(self, b: dict):
        s = _cf(b["samples"])
        bd = RefBurstData()
        bd.info = RefBurstInfo(b["id"], b["start"], b.get("stop", 0), b.get("last_active", 0),
                               b["center_bin"], b["magnitude"], b["noise"])
        bd.center_frequency = b["center_frequency"]
        bd.sample_rate = b["sample_rate"]
        bd.fft_size = b["fft_size"]
        bd.start_time_ns = b["start_time_ns"]
        bd.num_samples = s.shape[0]
        bd.samples = s.ctypes.data
        fr = C.POINTER(RefDownmixFrame)()
        n = self.L.burst_downmix_process(self.dm, C.byref(bd), C.byref(fr))
        if n <= 0 or not fr:
            return None
        f = fr.contents
        smp = np.ctypeslib.as_array(C.cast(f.samples, C.POINTER(C.c_float)),
                                    (2 * f.num_samples,)).copy().view(np.complex64)
        out = dict(id=f.id, timestamp=f.timestamp, center_frequency=f.center_frequency,
                   sample_rate=f.sample_rate, samples_per_symbol=f.samples_per_symbol,
                   direction=f.direction, magnitude=f.magnitude, noise=f.noise,
                   uw_start=f.uw_start, samples=smp)
        self.L.ref_free(f.samples)
        self.L.ref_free(fr)
        return out

    def demod(self, f: dict, gardner: bool = True):
        self.L.ref_set_gardner(int(gardner))
        s = _cf(f["samples"])
        df = RefDownmixFrame(f["id"], f["timestamp"], f["center_frequency"], f["sample_rate"],
                             f["samples_per_symbol"], f["direction"], f["magnitude"], f["noise"],
                             f["uw_start"], s.shape[0], s.ctypes.data)
        out = C.POINTER(RefDemodFrame)()
        ok = self.L.qpsk_demod(C.byref(df), C.byref(out))
        if not ok:
            return None
        d = out.contents
        bits = np.ctypeslib.as_array(d.bits, (d.n_bits,)).copy()
        llr = np.ctypeslib.as_array(d.llr, (d.n_bits,)).copy()
        res = dict(id=d.id, timestamp=d.timestamp, center_frequency=d.center_frequency,
                   direction=d.direction, magnitude=d.magnitude, noise=d.noise,
                   confidence=d.confidence, level=d.level, n_symbols=d.n_symbols,
                   n_payload_symbols=d.n_payload_symbols, bits=bits, llr=llr)
        self.L.ref_free(d.bits)
        self.L.ref_free(d.llr)
        self.L.ref_free(out)
        return res


def run_ref_binary(path: str, fmt: str = "cf32", sample_rate: int = 10_000_000,
                   center_freq: Optional[float] = None, extra: Optional[List[str]] = None,
                   file_info: str = "T"):
    """Run oracle/_ref/iridium-sniffer over a file; returns (stdout lines, stderr, wall s)."""
    import time
    cmd = [REF_BIN, "-f", path, f"--format={fmt}", "-r", str(sample_rate),
           f"--file-info={file_info}"]
    if center_freq is not None:
        cmd += ["-c", str(int(center_freq))]
    cmd += extra or []
    t = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True)
    return [l for l in r.stdout.splitlines() if l.startswith("RAW:")], r.stderr, time.time() - t


def parse_raw(line: str) -> dict:
    """Fields of a RAW: line (frame_output.c:182-192)."""
    f = line.split()
    mag_noise = f[4][2:]
    # N:%05.2f%+06.2f -> split at the sign of the second number
    k = max(mag_noise.rfind("+"), mag_noise.rfind("-"))
    return dict(file_info=f[1], ts_ms=float(f[2]), freq_hz=int(f[3]), magnitude=float(mag_noise[:k]),
                noise=float(mag_noise[k:]), id=int(f[5][2:]), confidence=int(f[6].rstrip("%")),
                level=float(f[7]), n_payload=int(f[8]), bits=f[9] if len(f) > 9 else "")
