"""The C-ABI library loads and exports every symbol include/iridium_b200.h declares (no GPU needed)."""
import importlib
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b((?:ir_|gpu_burst_fft_|burst_detector_|burst_downmix_|qpsk_demod|frame_decode|frame_output_|simd_init|ida_|gf2_remainder|bits_to_uint|uint_to_bits|bch_31_21_correct)\w*)\s*\(", txt)))


def test_library_exports_declared_symbols():
    pl = importlib.import_module("iridium-sniffer_b200.pipeline")
    if not os.path.exists(pl.LIB_PATH):
        pl.build_library()
    L = pl.load_library()
    names = []
    for h in os.listdir(os.path.join(ROOT, "include")):
        if h.endswith(".h"):
            names += _declared(h)
    assert "ir_pipeline_create" in names and "ir_pipeline_run_host" in names
    assert "frame_decode" in names and "ida_reassemble_flush" in names and "ir_classify_frames" in names
    names = [n for n in names if n != "ida_message_cb"]          # a typedef, not a function
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ but not exported"
    for n in pl.EXPORTED_SYMBOLS:
        assert n in names


def test_no_device_fails_loudly():
    """Without a CUDA device create must fail with a message, never fall back."""
    import ctypes as C
    pl = importlib.import_module("iridium-sniffer_b200.pipeline")
    L = pl.load_library()
    if L.ir_device_count() > 0:
        return
    try:
        pl.Pipeline()
    except RuntimeError as e:
        assert "no CUDA device" in str(e) or "CUDA" in str(e)
    else:
        raise AssertionError("Pipeline() succeeded without a GPU")


def test_classification_without_device_fails_loudly():
    """ir_classify_frames has no CPU path either: without a device it returns -1 and says why."""
    import ctypes as C
    import numpy as np
    pl = importlib.import_module("iridium-sniffer_b200.pipeline")
    L = pl.load_library()
    if L.ir_device_count() > 0:
        return
    try:
        pl.classify_frames([(np.zeros(100, np.uint8), None, 1)])
    except RuntimeError as e:
        assert "no CUDA device" in str(e)
    else:
        raise AssertionError("ir_classify_frames succeeded without a GPU")


def test_struct_sizes_match_header():
    import ctypes as C
    pl = importlib.import_module("iridium-sniffer_b200.pipeline")
    assert C.sizeof(pl.Frame) == 64 and C.sizeof(pl.Config) == 88
    assert C.sizeof(pl.Burst) == 104 and C.sizeof(pl.Results) == 112
    assert C.sizeof(pl.FrameClass) == 504
