"""SURVEY.md 8f rank 3, CPU side: the arithmetic of the device classification kernel
(iridium-sniffer_b200/csrc/frame_classify.cuh, compiled for the host by tests/fc_host_shim.cpp) against the
reference's own frame_decode() and ida_decode() (oracle/_ref/libref_frame.so, compiled unmodified) and the oracle
port, field for field, on generated IRA / IBC / IDA frames with and without bit errors, truncations and random
bits.  This pins WHAT the kernel computes; that the GPU computes the same is tests/test_zz_gpu_classify.py."""
import ctypes as C
import importlib.util
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PORT_SO = os.path.join(ROOT, "oracle", "libir_frame_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_frame.so")


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(HERE, name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


fg = _load("frame_gen")
fc = _load("frame_class_types")


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("fc") / "libfc_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-x", "c++",
                    os.path.join(HERE, "fc_host_shim.cpp"), "-o", out], check=True)
    lib = C.CDLL(out)
    assert lib.fc_host_sizeof() == C.sizeof(fc.FrameClass)
    lib.fc_host_classify.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(fc.FrameClass)]
    return lib


@pytest.fixture(scope="module")
def checkers():
    from oracle import bindings as ob
    if not os.path.exists(PORT_SO):
        ob.build(port=True, ref=False)
    libs = [("port", C.CDLL(PORT_SO), "orc_")]
    if os.path.exists(REF_SO) or os.path.exists("/root/reference/frame_decode.c"):
        if not os.path.exists(REF_SO):
            ob.build(port=False, ref=True)
        libs.append(("reference", C.CDLL(REF_SO), "ref_"))
    return [(n, fc.bind_checker(lib, pre)) for n, lib, pre in libs]


def _host(shim, bits, llr, direction):
    o = fc.FrameClass()
    lp = None if llr is None else llr.ctypes.data_as(C.c_void_p)
    shim.fc_host_classify(bits.ctypes.data_as(C.c_void_p), lp, len(bits), direction, C.byref(o))
    return o


def test_host_compiled_kernel_code_equals_reference_and_oracle(shim, checkers):
    cases = fg.corpus(101, 1600)
    n_type = {0: 0, 1: 0, 2: 0}
    n_ida = 0
    for bits, llr, direction in cases:
        got = _host(shim, bits, llr, direction)
        for name, chk in checkers:
            fc.assert_same(got, *chk(bits, llr, direction), where=name)
        n_type[got.frame_type] += 1
        n_ida += got.ida_ok
    # the corpus does exercise every outcome
    assert n_type[1] > 100 and n_type[2] > 100 and n_ida > 100 and n_type[0] > 400


def test_both_classifiers_run_on_every_frame(shim, checkers):
    """main.c:320-350 calls ida_decode() and frame_decode() independently: a frame can satisfy neither, and the
    two halves of the result must not disturb each other"""
    rng = np.random.default_rng(5)
    ida = np.array(fg.make_ida(rng, 9), np.uint8)
    o = _host(shim, ida, None, 1)
    assert o.ida_ok == 1 and o.da_len == 9 and o.crc_ok == 1 and o.payload_len == 9 and o.bch_len == 200
    ira = np.array(fg.make_ira(rng, 2), np.uint8)
    o = _host(shim, ira, None, 1)
    assert o.frame_type == 1 and o.n_pages == 2 and -90.0 <= o.lat <= 90.0
    for name, chk in checkers:
        fc.assert_same(o, *chk(ira, None, 1), where=name)
    empty = _host(shim, np.zeros(0, np.uint8), None, 1)
    assert bytes(empty) == bytes(fc.FrameClass())


def test_planted_frames_survive_the_whole_oracle_path(shim, checkers, port, synth):
    """The fixture the GPU end-to-end test relies on, proven here without a GPU: recordings whose bursts carry
    real IRA / IBC / IDA frames go through the CPU oracle (detect -> downmix -> demod), and the demodulated bits
    classify to the planted frames' fields -- by the oracle, the reference and the kernel's host-compiled code."""
    n_seen = {"ira": 0, "ibc": 0, "ida": 0}
    for rec, frames in fg.planted_recordings(synth):
        res, _ = port.run(rec.iq, center_frequency=rec.center_freq, sample_rate=rec.sample_rate)
        by_bits = {t.bits: t for t in rec.truth}
        matched = 0
        for r in res:
            bits = np.ascontiguousarray(r["bits"], np.uint8)
            text = "".join(map(str, bits))
            if text not in by_bits:
                continue
            matched += 1
            got = _host(shim, bits, None, r["direction"])
            want = _host(shim, np.array([int(c) for c in text], np.uint8), None, r["direction"])
            assert bytes(got) == bytes(want)
            for name, chk in checkers:
                fc.assert_same(got, *chk(bits, None, r["direction"]), where=name)
            assert got.frame_type != 0 or got.ida_ok == 1
            n_seen["ira"] += got.frame_type == 1
            n_seen["ibc"] += got.frame_type == 2
            n_seen["ida"] += got.ida_ok
        assert matched >= 6, (matched, len(rec.truth))
    assert n_seen["ira"] >= 3 and n_seen["ibc"] >= 2 and n_seen["ida"] >= 6, n_seen
