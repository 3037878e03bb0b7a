"""The GPU cases of the classification row (tests/gpu_classify_cases.py) had not run on a B200 when committed.  So that
their own code -- corpus, struct conversions, checkers, tolerances, count thresholds -- is not what fails there, two of
them run here with the library's classification entry points replaced by the SAME arithmetic compiled for the host
(tests/fc_host_shim.cpp: csrc/frame_classify.cuh, the code k_classify_frames runs one warp per frame) and, for the
pipeline, by the oracle-backed stand-ins of tests/blocks_host_shim.cpp.  Test infrastructure only; nothing here is a
statement about the GPU."""
import ctypes as C
import importlib
import importlib.util
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(HERE, name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.fixture(scope="module")
def host_classifier(tmp_path_factory):
    pl = importlib.import_module("iridium-sniffer_b200.pipeline")
    out = str(tmp_path_factory.mktemp("fc") / "libfc_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-x", "c++",
                    os.path.join(HERE, "fc_host_shim.cpp"), "-o", out], check=True)
    lib = C.CDLL(out)
    lib.fc_host_classify.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(pl.FrameClass)]

    def classify_frames(cases, device=0):
        res = []
        for bits, llr, direction in cases:
            bits = np.ascontiguousarray(bits, np.uint8)
            lp = None if llr is None else np.ascontiguousarray(llr, np.float32).ctypes.data_as(C.c_void_p)
            o = pl.FrameClass()
            lib.fc_host_classify(bits.ctypes.data_as(C.c_void_p), lp, len(bits), direction, C.byref(o))
            res.append(o)
        return res
    return pl, classify_frames


def test_dry_run_generated_frames(host_classifier, monkeypatch):
    pl, classify_frames = host_classifier
    cases = _load("gpu_classify_cases")
    monkeypatch.setattr(pl, "classify_frames", classify_frames)
    from oracle import bindings as ob
    if not os.path.exists(cases.PORT_SO):
        ob.build(port=True, ref=False)
    libs = [("port", cases.fc.bind_checker(C.CDLL(cases.PORT_SO), "orc_"))]
    if os.path.exists(cases.REF_SO):
        libs.append(("reference", cases.fc.bind_checker(C.CDLL(cases.REF_SO), "ref_")))
    cases.test_generated_frames_one_launch(pl, libs)


def test_dry_run_pipeline_classifies_planted_frames(host_classifier, monkeypatch, tmp_path, synth):
    """Pipeline.run_host + Pipeline.classify over the stand-ins (oracle path + the product's classification arithmetic
    on the host, no LLRs -- the planted frames arrive without bit errors), the case's own checks on top"""
    pl, classify_frames = host_classifier
    tb = _load("test_time_blocks")
    S = tb.build_blocks_shim(tmp_path)
    S.shim_set_devices(1, -1)
    cases = _load("gpu_classify_cases")
    monkeypatch.setattr(pl, "classify_frames", lambda cs, device=0: classify_frames([(b, None, d) for b, _, d in cs]))
    monkeypatch.setattr(pl, "Pipeline", tb.stand_in_pipeline_class(pl, S))
    chk = cases.fc.bind_checker(C.CDLL(cases.PORT_SO), "orc_")
    libs = [("port", lambda bits, llr, direction: chk(bits, None, direction))]
    cases.test_pipeline_classifies_planted_frames_from_device_memory(pl, libs, synth)


def test_dry_run_parsed_output_of_a_run(host_classifier, monkeypatch, tmp_path, synth):
    pl, _ = host_classifier
    tb = _load("test_time_blocks")
    S = tb.build_blocks_shim(tmp_path)
    S.shim_set_devices(1, -1)
    cases = _load("gpu_classify_cases")
    monkeypatch.setattr(pl, "Pipeline", tb.stand_in_pipeline_class(pl, S))
    chk = cases.fc.bind_checker(C.CDLL(cases.PORT_SO), "orc_")
    cases.test_parsed_output_of_a_run(pl, [("port", lambda bits, llr, direction: chk(bits, None, direction))], synth)


def test_dry_run_dropin_comparison_code(monkeypatch, tmp_path, synth):
    """tests/gpu_dropin_cases.py compares the linked drop-in program with the reference program; here the reference
    program stands on both sides, which runs the case's own parsing / sorting / tolerance code on real output"""
    cases = _load("gpu_dropin_cases")
    if not os.path.exists(cases.REF_BIN):
        pytest.skip("oracle/_ref/iridium-sniffer not built")
    monkeypatch.setattr(cases, "NEW_BIN", cases.REF_BIN)
    cases.test_reference_main_linked_against_the_library(synth, tmp_path)


def test_dry_run_reference_named_entry_points(host_classifier, monkeypatch, tmp_path):
    """frame_decode() / ida_decode() of csrc/refapi_frames.cu compiled for the host (tests/refapi_frames_host_shim.cpp),
    through the GPU case's own comparison with the reference's structs, byte for byte"""
    pl, _ = host_classifier
    cases = _load("gpu_classify_cases")
    out = str(tmp_path / "librf_shim.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-x", "c++",
                    os.path.join(HERE, "refapi_frames_host_shim.cpp"), "-o", out, "-L", os.path.dirname(pl.LIB_PATH),
                    "-l:libiridium_b200.so", "-Wl,-rpath," + os.path.dirname(pl.LIB_PATH)], check=True)
    S = C.CDLL(out)
    monkeypatch.setattr(pl, "load_library", lambda: S)
    monkeypatch.setattr(cases, "GEO_TOL", 0.0)          # host-compiled: the C library's atan2 on both sides
    cases.test_reference_named_entry_points(pl)
