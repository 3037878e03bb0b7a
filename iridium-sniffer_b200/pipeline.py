"""ctypes mirror of include/iridium_b200.h -- the host-side handle the tests and bench.py use.

The arithmetic lives entirely in libiridium_b200.so (hand-written sm_100a kernels).  There
is no Python / NumPy / oracle fallback: if the library or a CUDA device is missing, loading
or creating a pipeline raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libiridium_b200.so")
CSRC = os.path.join(HERE, "csrc")

FMT_CF32, FMT_CI16, FMT_CI8 = 0, 1, 2
FMT_BY_NAME = {"cf32": FMT_CF32, "ci16": FMT_CI16, "ci8": FMT_CI8}
ABI_VERSION = 1


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("device", C.c_int32),
                ("center_frequency", C.c_double), ("sample_rate", C.c_int32),
                ("fft_size", C.c_int32), ("burst_width_hz", C.c_int32),
                ("threshold_db", C.c_float), ("use_gardner", C.c_int32),
                ("feed_block", C.c_int32), ("start_time_ns", C.c_uint64),
                ("max_samples", C.c_uint64), ("h2d_chunk", C.c_int32),
                ("reserved", C.c_int32 * 7)]


class Frame(C.Structure):
    _fields_ = [("id", C.c_uint64), ("timestamp", C.c_uint64), ("center_frequency", C.c_double),
                ("direction", C.c_int32), ("magnitude", C.c_float), ("noise", C.c_float),
                ("confidence", C.c_int32), ("level", C.c_float), ("n_symbols", C.c_int32),
                ("n_payload_symbols", C.c_int32), ("n_bits", C.c_int32),
                ("bits_offset", C.c_uint32)]


class Burst(C.Structure):
    _fields_ = [("id", C.c_uint64), ("start", C.c_uint64), ("stop", C.c_uint64),
                ("last_active", C.c_uint64), ("center_bin", C.c_int32), ("magnitude", C.c_float),
                ("noise", C.c_float), ("num_samples", C.c_uint64), ("emit_count", C.c_uint64),
                ("downmix_status", C.c_int32), ("demod_ok", C.c_int32),
                ("center_offset", C.c_float), ("dm_start", C.c_int32), ("uw_start", C.c_int32),
                ("frame_len", C.c_int32), ("uw_start_frac", C.c_float),
                ("dm_direction", C.c_int32), ("dec_len", C.c_int32)]


class FrameClass(C.Structure):
    """ir_frame_class_t: frame_decode() + ida_decode() outcome of one frame"""
    _fields_ = [("frame_type", C.c_int32), ("sat_id", C.c_int32), ("beam_id", C.c_int32), ("lat", C.c_double),
                ("lon", C.c_double), ("alt", C.c_int32), ("pos_xyz", C.c_int32 * 3), ("n_pages", C.c_int32),
                ("tmsi", C.c_uint32 * 12), ("msc_id", C.c_int32 * 12), ("timeslot", C.c_int32),
                ("sv_blocking", C.c_int32), ("bc_type", C.c_int32), ("iri_time", C.c_uint32), ("ida_ok", C.c_int32),
                ("lcw_ft", C.c_int32), ("lcw_code", C.c_int32), ("ec_lcw", C.c_int32), ("lcw3_val", C.c_uint32),
                ("da_ctr", C.c_int32), ("da_len", C.c_int32), ("cont", C.c_int32), ("payload_len", C.c_int32),
                ("crc_ok", C.c_int32), ("fixederrs", C.c_int32), ("bch_len", C.c_int32), ("stored_crc", C.c_uint16),
                ("computed_crc", C.c_uint16), ("payload", C.c_uint8 * 32), ("bch_stream", C.c_uint8 * 256)]


class Block(C.Structure):
    """ir_block_t: one time block of a long stream (feed range and owned range, in samples)"""
    _fields_ = [("feed_first", C.c_uint64), ("feed_end", C.c_uint64), ("own_first", C.c_uint64), ("own_end", C.c_uint64)]


BLOCK_ID_STRIDE = 1_000_000_000


class MultiResults(C.Structure):
    """ir_multi_results_t"""
    _fields_ = [("n_frames", C.c_size_t), ("frames", C.POINTER(Frame)), ("block", C.POINTER(C.c_uint32)),
                ("index", C.POINTER(C.c_uint32)), ("classes", C.POINTER(C.POINTER(FrameClass))),
                ("n_blocks", C.c_size_t), ("blocks", C.POINTER(Block)), ("bits", C.POINTER(C.POINTER(C.c_uint8))),
                ("llr", C.POINTER(C.POINTER(C.c_float))), ("start_time_ns", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("samples_fed", C.c_uint64)]


class Results(C.Structure):
    _fields_ = [("n_bursts", C.c_size_t), ("bursts", C.POINTER(Burst)),
                ("n_frames", C.c_size_t), ("frames", C.POINTER(Frame)),
                ("bits", C.POINTER(C.c_uint8)), ("llr", C.POINTER(C.c_float)),
                ("n_bits_total", C.c_size_t), ("ms_total", C.c_float),
                ("ms_detect_fft", C.c_float), ("ms_detect_scan", C.c_float),
                ("ms_downmix_fir", C.c_float), ("ms_downmix_chain", C.c_float),
                ("ms_demod", C.c_float), ("kernel_launches", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("alg_bytes", C.c_uint64)]


def build_library(verbose: bool = False) -> str:
    """Compile csrc/ for sm_100a with nvcc (cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j4"], capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libiridium_b200.so failed:\n" + (r.stdout or "") + (r.stderr or ""))
    return LIB_PATH


_lib = None


def load_library() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run __graft_entry__.build() (there is no fallback path)")
    L = C.CDLL(LIB_PATH)
    L.ir_last_error.restype = C.c_char_p
    L.ir_device_count.restype = C.c_int
    L.ir_pipeline_create.restype = C.c_void_p
    L.ir_pipeline_create.argtypes = [C.POINTER(Config)]
    L.ir_pipeline_destroy.argtypes = [C.c_void_p]
    L.ir_pipeline_reset.argtypes = [C.c_void_p]
    L.ir_pipeline_run_host.restype = C.c_int
    L.ir_pipeline_run_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    L.ir_pipeline_run_device.restype = C.c_int
    L.ir_pipeline_run_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    L.ir_pipeline_results.restype = C.c_int
    L.ir_pipeline_results.argtypes = [C.c_void_p, C.POINTER(Results)]
    L.ir_pipeline_copy_mag.restype = C.c_int
    L.ir_pipeline_copy_mag.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
    for name in ("ir_pipeline_copy_frame_samples", "ir_pipeline_copy_decimated",
                 "ir_pipeline_copy_burst_samples"):
        f = getattr(L, name)
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    L.ir_format_raw.restype = C.c_int
    L.ir_format_raw.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_uint64, C.POINTER(Frame),
                                C.c_void_p]
    L.ir_pipeline_format_raw_all.restype = C.c_long
    L.ir_pipeline_format_raw_all.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_char_p, C.c_size_t]
    L.ir_host_alloc.restype = C.c_void_p
    L.ir_host_alloc.argtypes = [C.c_size_t]
    L.ir_host_free.argtypes = [C.c_void_p]
    L.ir_plan_chunks.restype = C.c_long
    L.ir_plan_chunks.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, C.POINTER(C.c_size_t), C.c_size_t]
    L.ir_format_fixed.restype = C.c_int
    L.ir_format_fixed.argtypes = [C.c_char_p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int]
    L.ir_pipeline_scan_stats.restype = C.c_int
    L.ir_pipeline_scan_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_int]
    L.ir_classify_frames.restype = C.c_int
    L.ir_classify_frames.argtypes = [C.c_int, C.POINTER(Frame), C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                                     C.POINTER(FrameClass)]
    L.ir_pipeline_classify.restype = C.c_long
    L.ir_pipeline_classify.argtypes = [C.c_void_p, C.POINTER(FrameClass), C.c_size_t]
    L.ir_pipeline_last_classify_ms.restype = C.c_float
    L.ir_pipeline_last_classify_ms.argtypes = [C.c_void_p]
    L.ir_format_lcw.restype = C.c_int
    L.ir_format_lcw.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p]
    L.ir_format_ida.restype = C.c_int
    L.ir_format_ida.argtypes = [C.c_char_p, C.c_size_t, C.c_uint64, C.POINTER(Frame), C.c_void_p]
    L.ir_pipeline_format_parsed_all.restype = C.c_long
    L.ir_pipeline_format_parsed_all.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_void_p, C.c_size_t, C.c_char_p,
                                                C.c_size_t]
    L.ir_block_halo.restype = C.c_size_t
    L.ir_block_halo.argtypes = [C.POINTER(Config)]
    L.ir_block_tail.restype = C.c_size_t
    L.ir_block_tail.argtypes = [C.POINTER(Config)]
    L.ir_plan_blocks.restype = C.c_long
    L.ir_plan_blocks.argtypes = [C.POINTER(Config), C.c_size_t, C.c_int, C.POINTER(Block), C.c_size_t]
    L.ir_pipeline_set_origin.restype = C.c_int
    L.ir_pipeline_set_origin.argtypes = [C.c_void_p, C.c_uint64]
    L.ir_merge_blocks.restype = C.c_long
    L.ir_merge_blocks.argtypes = [C.POINTER(Config), C.c_uint64, C.POINTER(Block), C.c_int, C.POINTER(C.POINTER(Frame)),
                                  C.POINTER(C.c_size_t), C.POINTER(Frame), C.POINTER(C.c_uint32), C.c_size_t]
    L.ir_pipeline_set_start_time.restype = C.c_int
    L.ir_pipeline_set_start_time.argtypes = [C.c_void_p, C.c_uint64]
    L.ir_multi_create.restype = C.c_void_p
    L.ir_multi_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_int), C.c_int]
    L.ir_multi_destroy.argtypes = [C.c_void_p]
    L.ir_multi_run_host.restype = C.c_int
    L.ir_multi_run_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
    L.ir_multi_results.restype = C.c_int
    L.ir_multi_results.argtypes = [C.c_void_p, C.POINTER(MultiResults)]
    L.ir_multi_format_raw_all.restype = C.c_long
    L.ir_multi_format_raw_all.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_char_p, C.c_size_t]
    L.ir_multi_run_streams_host.restype = C.c_int
    L.ir_multi_run_streams_host.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_int, C.c_int]
    L.ir_multi_set_classify.restype = C.c_int
    L.ir_multi_set_classify.argtypes = [C.c_void_p, C.c_int]
    L.ir_multi_format_parsed_all.restype = C.c_long
    L.ir_multi_format_parsed_all.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_char_p, C.c_size_t]
    _lib = L
    return L


def make_config(sample_rate: int = 10_000_000, center_frequency: float = 1_622_000_000.0, device: int = 0,
                threshold_db: float = 16.0, use_gardner: bool = True, fft_size: int = 0, feed_block: int = 32768,
                start_time_ns: int = 0, h2d_chunk: int = 0) -> Config:
    cfg = Config()
    cfg.abi_version = ABI_VERSION
    cfg.device = device
    cfg.center_frequency = center_frequency
    cfg.sample_rate = sample_rate
    cfg.fft_size = fft_size
    cfg.threshold_db = threshold_db
    cfg.use_gardner = int(use_gardner)
    cfg.feed_block = feed_block
    cfg.start_time_ns = start_time_ns
    cfg.h2d_chunk = h2d_chunk
    return cfg


# ---- one long stream over several pipelines by time blocks: plan and merge, host bookkeeping
def plan_blocks(cfg: Config, n_samples: int, n_blocks: int) -> List[Block]:
    L = load_library()
    out = (Block * max(n_blocks, 1))()
    k = L.ir_plan_blocks(C.byref(cfg), n_samples, n_blocks, out, max(n_blocks, 1))
    if k < 0:
        raise RuntimeError("ir_plan_blocks failed: " + L.ir_last_error().decode())
    res = []
    for i in range(k):
        b = Block()
        C.memmove(C.byref(b), C.byref(out[i]), C.sizeof(Block))
        res.append(b)
    return res


def merge_blocks(cfg: Config, start_time_ns: int, blocks: List[Block], frame_lists: List[List[dict]]) -> List[dict]:
    """ir_merge_blocks over per-block frame lists (dicts with the ir_frame_t fields, as RunResult.frames holds them;
    other keys -- bits, llr -- ride along).  Returns the kept frames in time order, each a copy with the merged id and
    a "block" key."""
    L = load_library()
    nb = len(blocks)
    assert len(frame_lists) == nb
    names = [k for k, _ in Frame._fields_]
    arrays, ptrs, counts = [], (C.POINTER(Frame) * max(nb, 1))(), (C.c_size_t * max(nb, 1))()
    for k, fl in enumerate(frame_lists):
        a = (Frame * max(len(fl), 1))()
        for i, d in enumerate(fl):
            for nm in names:
                if nm in d:
                    setattr(a[i], nm, d[nm])
            a[i].bits_offset = i                       # carries the list index through the merge
        arrays.append(a)
        ptrs[k] = C.cast(a, C.POINTER(Frame))
        counts[k] = len(fl)
    cap = sum(len(fl) for fl in frame_lists)
    out, out_block = (Frame * max(cap, 1))(), (C.c_uint32 * max(cap, 1))()
    barr = (Block * max(nb, 1))(*blocks)
    n = L.ir_merge_blocks(C.byref(cfg), start_time_ns, barr, nb, ptrs, counts, out, out_block, cap)
    if n < 0:
        raise RuntimeError("ir_merge_blocks failed: " + L.ir_last_error().decode())
    res = []
    for i in range(n):
        d = dict(frame_lists[out_block[i]][out[i].bits_offset])
        d["id"] = out[i].id
        d["block"] = int(out_block[i])
        res.append(d)
    return res


def classify_frames(cases, device: int = 0):
    """ir_classify_frames over [(bits uint8[n], llr float32[n] or None, direction)]: one launch for all of them.
    Either every case carries LLRs or none does (the C call takes one llr array or NULL)."""
    L = load_library()
    n = len(cases)
    with_llr = n > 0 and cases[0][1] is not None
    if any((c[1] is not None) != with_llr for c in cases):
        raise ValueError("mixed LLR / no-LLR cases: call once per kind")
    frames = (Frame * max(n, 1))()
    off = 0
    for i, (bits, llr, direction) in enumerate(cases):
        frames[i].n_bits, frames[i].direction, frames[i].bits_offset = len(bits), direction, off
        off += len(bits)
    allb = np.concatenate([np.asarray(c[0], np.uint8) for c in cases]) if n else np.zeros(0, np.uint8)
    allb = np.ascontiguousarray(np.append(allb, np.uint8(0)))
    alll = None
    if with_llr:
        alll = np.ascontiguousarray(np.append(np.concatenate([np.asarray(c[1], np.float32) for c in cases]), np.float32(0)))
    out = (FrameClass * max(n, 1))()
    rc = L.ir_classify_frames(device, frames, n, allb.ctypes.data_as(C.c_void_p),
                              alll.ctypes.data_as(C.c_void_p) if with_llr else None, off, out)
    if rc != 0:
        raise RuntimeError("ir_classify_frames failed: " + L.ir_last_error().decode())
    return [out[i] for i in range(n)]


EXPORTED_SYMBOLS = [
    "ir_last_error", "ir_device_count", "ir_pipeline_create", "ir_pipeline_destroy",
    "ir_pipeline_reset", "ir_pipeline_run_host", "ir_pipeline_run_device", "ir_pipeline_results",
    "ir_pipeline_copy_mag", "ir_pipeline_copy_frame_samples", "ir_pipeline_copy_decimated",
    "ir_pipeline_copy_burst_samples", "ir_format_raw", "ir_pipeline_format_raw_all", "ir_host_alloc",
    "ir_host_free", "ir_format_fixed", "ir_pipeline_scan_stats", "ir_plan_chunks", "ir_classify_frames", "ir_pipeline_classify",
    "ir_format_lcw", "ir_format_ida", "ir_pipeline_format_parsed_all", "ir_pipeline_last_classify_ms",
    "ir_fill_decoded_frame", "ir_fill_ida_burst",
    "ir_block_halo", "ir_block_tail", "ir_plan_blocks", "ir_pipeline_set_origin", "ir_merge_blocks",
    "ir_pipeline_set_start_time", "ir_multi_create", "ir_multi_destroy", "ir_multi_run_host", "ir_multi_results",
    "ir_multi_format_raw_all", "ir_multi_set_classify", "ir_multi_format_parsed_all", "ir_multi_run_streams_host",
]


class RunResult:
    """Python view of ir_results_t (copies everything out of the library's buffers)."""

    def __init__(self, r: Results):
        self.stats = {k: getattr(r, k) for k in (
            "ms_total", "ms_detect_fft", "ms_detect_scan", "ms_downmix_fir", "ms_downmix_chain",
            "ms_demod", "kernel_launches", "h2d_bytes", "d2h_bytes", "alg_bytes")}
        self.bursts: List[dict] = []
        for i in range(r.n_bursts):
            b = r.bursts[i]
            self.bursts.append({k: getattr(b, k) for k, _ in Burst._fields_})
        allbits = np.ctypeslib.as_array(r.bits, (max(r.n_bits_total, 1),)).copy() if r.n_bits_total else np.zeros(0, np.uint8)
        allllr = np.ctypeslib.as_array(r.llr, (max(r.n_bits_total, 1),)).copy() if r.n_bits_total else np.zeros(0, np.float32)
        self.frames: List[dict] = []
        self._cframes = []
        for i in range(r.n_frames):
            f = r.frames[i]
            d = {k: getattr(f, k) for k, _ in Frame._fields_}
            d["bits"] = allbits[f.bits_offset:f.bits_offset + f.n_bits]
            d["llr"] = allllr[f.bits_offset:f.bits_offset + f.n_bits]
            self.frames.append(d)
            cf = Frame()
            C.memmove(C.byref(cf), C.byref(f), C.sizeof(Frame))
            self._cframes.append(cf)

    def raw_lines(self, file_info: str = "T", t0: Optional[int] = None) -> List[str]:
        """RAW: lines as frame_output_print would emit them (frame_output.c:144-199)."""
        L = load_library()
        if not self.frames:
            return []
        if t0 is None:      # ensure_initialized: first printed frame's timestamp floored to 1 s
            t0 = (self.frames[0]["timestamp"] // 1_000_000_000) * 1_000_000_000
        out = []
        buf = C.create_string_buffer(4096)
        for cf, d in zip(self._cframes, self.frames):
            bits = np.ascontiguousarray(d["bits"], np.uint8)
            n = L.ir_format_raw(buf, 4096, file_info.encode(), t0, C.byref(cf),
                                bits.ctypes.data_as(C.c_void_p))
            if n < 0:
                raise RuntimeError("ir_format_raw failed")
            out.append(buf.value.decode())
        return out


class Multi:
    """ir_multi_t: one stream in time blocks over several GPUs of one process (a pipeline and a thread per device)."""

    def __init__(self, devices, **cfg_kw):
        self.L = load_library()
        self.cfg = make_config(**cfg_kw)
        devs = (C.c_int * len(devices))(*devices)
        self.h = self.L.ir_multi_create(C.byref(self.cfg), devs, len(devices))
        if not self.h:
            raise RuntimeError("ir_multi_create failed: " + self.L.ir_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.ir_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run_host(self, iq: np.ndarray, fmt: str = "cf32", n_blocks: int = 0) -> List[dict]:
        a, n = Pipeline._as_raw(iq, fmt)
        if self.L.ir_multi_run_host(self.h, a.ctypes.data_as(C.c_void_p), n, FMT_BY_NAME[fmt], n_blocks) != 0:
            raise RuntimeError("ir_multi_run_host failed: " + self.L.ir_last_error().decode())
        return self.frames()

    def run_streams_host(self, streams, fmt: str = "cf32") -> List[dict]:
        """independent streams (one numpy array each), stream s on device s % n_devices; frames of all of them back,
        `block` = the stream's index"""
        arrs = [Pipeline._as_raw(a, fmt) for a in streams]
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a, _ in arrs])
        ns = (C.c_size_t * len(arrs))(*[n for _, n in arrs])
        if self.L.ir_multi_run_streams_host(self.h, ptrs, ns, len(arrs), FMT_BY_NAME[fmt]) != 0:
            raise RuntimeError("ir_multi_run_streams_host failed: " + self.L.ir_last_error().decode())
        return self.frames()

    def results(self) -> MultiResults:
        r = MultiResults()
        if self.L.ir_multi_results(self.h, C.byref(r)) != 0:
            raise RuntimeError("ir_multi_results failed: " + self.L.ir_last_error().decode())
        return r

    def frames(self) -> List[dict]:
        r = self.results()
        out = []
        for i in range(r.n_frames):
            f = r.frames[i]
            d = {k: getattr(f, k) for k, _ in Frame._fields_}
            d["block"] = int(r.block[i])
            d["bits"] = np.ctypeslib.as_array(r.bits[d["block"]], (f.bits_offset + f.n_bits,))[f.bits_offset:].copy()
            out.append(d)
        return out

    def _text(self, fn, what: str, file_info: str, t0: int) -> bytes:
        need = fn(self.h, file_info.encode(), t0, None, 0)
        if need < 0:
            raise RuntimeError(what + " failed: " + self.L.ir_last_error().decode())
        buf = C.create_string_buffer(max(need, 1))
        n = fn(self.h, file_info.encode(), t0, buf, need)
        if n < 0:
            raise RuntimeError(what + " failed: " + self.L.ir_last_error().decode())
        return buf.raw[:n]

    def raw_text(self, file_info: str = "T", t0: int = 0) -> bytes:
        return self._text(self.L.ir_multi_format_raw_all, "ir_multi_format_raw_all", file_info, t0)

    def set_classify(self, on: bool = True) -> None:
        if self.L.ir_multi_set_classify(self.h, int(on)) != 0:
            raise RuntimeError("ir_multi_set_classify failed: " + self.L.ir_last_error().decode())

    def parsed_text(self, file_info: str = "T", t0: int = 0) -> bytes:
        """the `--parsed` text of the merged run (needs set_classify() before the run)"""
        return self._text(self.L.ir_multi_format_parsed_all, "ir_multi_format_parsed_all", file_info, t0)


class Pipeline:
    """detect -> downmix -> demod on one B200 (ir_pipeline_t)."""

    def __init__(self, sample_rate: int = 10_000_000, center_frequency: float = 1_622_000_000.0,
                 device: int = 0, threshold_db: float = 16.0, use_gardner: bool = True,
                 fft_size: int = 0, feed_block: int = 32768, start_time_ns: int = 0,
                 h2d_chunk: int = 0):
        self.L = load_library()
        cfg = make_config(sample_rate, center_frequency, device, threshold_db, use_gardner, fft_size, feed_block,
                          start_time_ns, h2d_chunk)
        self.cfg = cfg
        self.h = self.L.ir_pipeline_create(C.byref(cfg))
        if not self.h:
            raise RuntimeError("ir_pipeline_create failed: " + self.L.ir_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.ir_pipeline_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise RuntimeError(f"{what} failed: " + self.L.ir_last_error().decode())

    @staticmethod
    def _as_raw(iq: np.ndarray, fmt: str):
        if fmt == "cf32":
            a = np.ascontiguousarray(iq, np.complex64)
            return a, a.shape[0]
        if fmt == "ci16":
            a = np.ascontiguousarray(iq, np.int16)
            return a, a.shape[0] // 2
        a = np.ascontiguousarray(iq, np.int8)
        return a, a.shape[0] // 2

    def run_host(self, iq: np.ndarray, fmt: str = "cf32") -> RunResult:
        a, n = self._as_raw(iq, fmt)
        self._check(self.L.ir_pipeline_run_host(self.h, a.ctypes.data_as(C.c_void_p), n,
                                                FMT_BY_NAME[fmt]), "ir_pipeline_run_host")
        return self.results()

    def run_host_ptr(self, ptr: int, n_samples: int, fmt: str = "cf32") -> RunResult:
        self._check(self.L.ir_pipeline_run_host(self.h, C.c_void_p(ptr), n_samples,
                                                FMT_BY_NAME[fmt]), "ir_pipeline_run_host")
        return self.results()

    def run_device_ptr(self, dev_ptr: int, n_samples: int, fmt: str = "cf32") -> RunResult:
        self._check(self.L.ir_pipeline_run_device(self.h, C.c_void_p(dev_ptr), n_samples,
                                                  FMT_BY_NAME[fmt]), "ir_pipeline_run_device")
        return self.results()

    def set_origin(self, sample_origin: int) -> None:
        """ir_pipeline_set_origin: absolute index of sample 0 of the following runs (time stamps only)."""
        self._check(self.L.ir_pipeline_set_origin(self.h, sample_origin), "ir_pipeline_set_origin")

    def run_block(self, iq: np.ndarray, block: Block) -> RunResult:
        """One time block of the cf32 stream `iq` (the whole stream, host memory): feed range in, frames stamped
        on the stream's clock out.  The caller merges the blocks' results with merge_blocks()."""
        a = np.ascontiguousarray(iq[block.feed_first:block.feed_end], np.complex64)
        self.set_origin(block.feed_first)
        try:
            return self.run_host(a, "cf32")
        finally:
            self.set_origin(0)

    # ---- bare calls (no Python-side result conversion): what bench.py times
    def run_device_raw(self, dev_ptr: int, n_samples: int, fmt: str = "cf32") -> None:
        self._check(self.L.ir_pipeline_run_device(self.h, C.c_void_p(dev_ptr), n_samples,
                                                  FMT_BY_NAME[fmt]), "ir_pipeline_run_device")

    def run_host_raw(self, ptr: int, n_samples: int, fmt: str = "cf32") -> None:
        self._check(self.L.ir_pipeline_run_host(self.h, C.c_void_p(ptr), n_samples,
                                                FMT_BY_NAME[fmt]), "ir_pipeline_run_host")

    def raw_text(self, file_info: str = "T", t0: int = 0) -> bytes:
        """Every RAW: line of the last run, formatted inside the library in one call."""
        need = self.L.ir_pipeline_format_raw_all(self.h, file_info.encode(), t0, None, 0)
        if need < 0:
            raise RuntimeError("ir_pipeline_format_raw_all failed")
        if getattr(self, "_txt_cap", 0) < need:
            self._txt = C.create_string_buffer(need)
            self._txt_cap = need
        n = self.L.ir_pipeline_format_raw_all(self.h, file_info.encode(), t0, self._txt, self._txt_cap)
        if n < 0:
            raise RuntimeError("ir_pipeline_format_raw_all failed: " + self.L.ir_last_error().decode())
        return self._txt.raw[:n]

    def raw_text_len(self, file_info: str = "T", t0: int = 0) -> int:
        """Bare C call of the batched sink into a buffer kept by this object (what bench.py times);
        returns the number of bytes.  `raw_text_view()` exposes them without another copy."""
        fi = file_info.encode()
        if getattr(self, "_txt_cap", 0) == 0:
            need = self.L.ir_pipeline_format_raw_all(self.h, fi, t0, None, 0)
            if need < 0:
                raise RuntimeError("ir_pipeline_format_raw_all failed")
            self._txt_cap = max(int(need) * 2, 1 << 20)
            self._txt = C.create_string_buffer(self._txt_cap)
        n = self.L.ir_pipeline_format_raw_all(self.h, fi, t0, self._txt, self._txt_cap)
        if n < 0:                                     # grown since the buffer was sized: size it again
            need = self.L.ir_pipeline_format_raw_all(self.h, fi, t0, None, 0)
            if need < 0:
                raise RuntimeError("ir_pipeline_format_raw_all failed: " + self.L.ir_last_error().decode())
            self._txt_cap = int(need) * 2
            self._txt = C.create_string_buffer(self._txt_cap)
            n = self.L.ir_pipeline_format_raw_all(self.h, fi, t0, self._txt, self._txt_cap)
            if n < 0:
                raise RuntimeError("ir_pipeline_format_raw_all failed: " + self.L.ir_last_error().decode())
        self._txt_len = int(n)
        return int(n)

    def raw_text_view(self) -> memoryview:
        return memoryview(self._txt)[:getattr(self, "_txt_len", 0)]

    def classify(self) -> list:
        """frame_decode() + ida_decode() outcome of every frame of the last run (one launch, bits / LLRs read
        where the demod kernel left them)"""
        r = Results()
        self._check(self.L.ir_pipeline_results(self.h, C.byref(r)), "ir_pipeline_results")
        n = int(r.n_frames)
        out = (FrameClass * max(n, 1))()
        got = self.L.ir_pipeline_classify(self.h, out, n)
        if got < 0:
            raise RuntimeError("ir_pipeline_classify failed: " + self.L.ir_last_error().decode())
        return [out[i] for i in range(got)]

    def classify_ms(self) -> float:
        return float(self.L.ir_pipeline_last_classify_ms(self.h))

    def parsed_text(self, file_info: str = "T", t0: int = 0) -> bytes:
        """The run's output the way `--parsed` prints it: IDA lines where ida_decode() accepts, RAW lines otherwise
        (classification on the GPU inside the call)."""
        need = self.L.ir_pipeline_format_parsed_all(self.h, file_info.encode(), t0, None, 0, None, 0)
        if need < 0:
            raise RuntimeError("ir_pipeline_format_parsed_all failed: " + self.L.ir_last_error().decode())
        buf = C.create_string_buffer(int(need) + 1)
        n = self.L.ir_pipeline_format_parsed_all(self.h, file_info.encode(), t0, None, 0, buf, len(buf))
        if n < 0:
            raise RuntimeError("ir_pipeline_format_parsed_all failed: " + self.L.ir_last_error().decode())
        return buf.raw[:n]

    def scan_stats(self) -> dict:
        """Counters of the detector state machine over the last run (ir_pipeline_scan_stats)."""
        a = (C.c_uint64 * 8)()
        rc = self.L.ir_pipeline_scan_stats(self.h, a, 8)
        if rc < 0:
            raise RuntimeError("ir_pipeline_scan_stats failed")
        keys = ("launches_kept", "launches_bailed", "commands", "event_frames", "exact_words", "waits", "last_bail_frame")
        d = {k: int(a[i]) for i, k in enumerate(keys)}
        d["bitmap_rebuilds"] = int(a[7])
        d["streaming"] = rc == 1
        d["segmented"] = rc == 2          # launches_* count chunks, "commands" counts rounds
        if rc == 2:
            v = d.pop("last_bail_frame")
            d["last_bail_reason"] = v & 0xFF
            d["generic_segment_walks"] = v >> 8      # segments with more than 32 bursts alive (all rounds)
        return d

    def stats(self) -> dict:
        r = Results()
        self._check(self.L.ir_pipeline_results(self.h, C.byref(r)), "ir_pipeline_results")
        d = {k: getattr(r, k) for k in ("ms_total", "ms_detect_fft", "ms_detect_scan", "ms_downmix_fir",
                                        "ms_downmix_chain", "ms_demod", "kernel_launches", "h2d_bytes",
                                        "d2h_bytes", "alg_bytes")}
        d["n_bursts"], d["n_frames"] = r.n_bursts, r.n_frames
        return d

    def results(self) -> RunResult:
        r = Results()
        self._check(self.L.ir_pipeline_results(self.h, C.byref(r)), "ir_pipeline_results")
        return RunResult(r)

    # ---- parity taps
    def mag(self, frame0: int, n_frames: int, fft_size: int) -> np.ndarray:
        out = np.empty((n_frames, fft_size), np.float32)
        self._check(self.L.ir_pipeline_copy_mag(self.h, frame0, n_frames,
                                                out.ctypes.data_as(C.c_void_p)), "ir_pipeline_copy_mag")
        return out

    def _copy_cf(self, fn, index: int, cap: int) -> Optional[np.ndarray]:
        out = np.empty(cap, np.complex64)
        n = fn(self.h, index, out.ctypes.data_as(C.c_void_p), cap)
        return None if n < 0 else out[:n].copy()

    def frame_samples(self, burst_index: int) -> Optional[np.ndarray]:
        return self._copy_cf(self.L.ir_pipeline_copy_frame_samples, burst_index, 4440)

    def decimated(self, burst_index: int, cap: int = 60000) -> Optional[np.ndarray]:
        return self._copy_cf(self.L.ir_pipeline_copy_decimated, burst_index, cap)

    def burst_samples(self, burst_index: int, cap: int = 2 * 1024 * 1024 + 16) -> Optional[np.ndarray]:
        return self._copy_cf(self.L.ir_pipeline_copy_burst_samples, burst_index, cap)
