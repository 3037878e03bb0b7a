"""The drop-in claim as a link (no GPU needed): the reference's own main.c, options.c, gsmtap.c, web_map.c,
doppler_pos.c and sbd_acars.c -- compiled unmodified -- link against libiridium_b200.so IN PLACE OF burst_detect.c,
burst_downmix.c, qpsk_demod.c, fir_filter.c, window_func.c, simd_*.c, frame_decode.c, ida_decode.c and
frame_output.c (oracle/Makefile target `dropin`), and the library's weak references bind to main.c's queues,
counters and switches.  That the linked program then prints what the reference prints is the GPU case in
tests/gpu_dropin_cases.py."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "iridium-sniffer-b200")


@pytest.fixture(scope="module")
def dropin():
    if os.path.exists("/root/reference/main.c"):
        import importlib
        pl = importlib.import_module("iridium-sniffer_b200.pipeline")
        if not os.path.exists(pl.LIB_PATH):
            pl.build_library()
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref", "dropin"], check=True)
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/iridium-sniffer-b200 not built and /root/reference absent")
    return BIN


def test_reference_main_links_against_the_library(dropin):
    needed = subprocess.run(["readelf", "-d", dropin], capture_output=True, text=True, check=True).stdout
    assert "libiridium_b200.so" in needed and "fftw" not in needed
    undefined = subprocess.run(["nm", "-D", "--undefined-only", dropin], capture_output=True, text=True, check=True).stdout
    for sym in ("burst_detector_create", "burst_detector_thread", "burst_downmix_create", "burst_downmix_thread", "qpsk_demod",
                "frame_decode", "ida_decode", "ida_reassemble", "frame_output_print", "frame_output_print_ida", "simd_init"):
        assert sym in undefined, sym                 # main.c takes every stage from the library
    exported = subprocess.run(["nm", "-D", "--defined-only", dropin], capture_output=True, text=True, check=True).stdout
    for sym in ("samples_queue", "burst_queue", "frame_queue", "blocking_queue_take", "blocking_queue_put", "blocking_queue_add",
                "stat_n_detected", "stat_n_dropped", "use_gardner", "verbose", "diagnostic_mode", "acars_enabled"):
        assert sym in exported, sym                  # ... and the library's weak references find main.c's definitions
    r = subprocess.run([dropin, "--help"], capture_output=True, text=True)
    assert "Usage: iridium-sniffer" in r.stdout + r.stderr


def test_linked_program_without_a_gpu_says_so(dropin, tmp_path):
    import ctypes as C
    import importlib
    import numpy as np
    L = importlib.import_module("iridium-sniffer_b200.pipeline").load_library()
    if L.ir_device_count() > 0:
        pytest.skip("a GPU is present: covered by tests/gpu_dropin_cases.py")
    path = str(tmp_path / "noise.cf32")
    (np.random.default_rng(0).standard_normal(2 * 1_000_000).astype(np.float32) * 0.01).tofile(path)
    r = subprocess.run([dropin, "-f", path, "--format=cf32", "-r", "10000000", "-c", "1622000000"], capture_output=True,
                       text=True, timeout=120)
    assert "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr and r.stdout == ""
