"""Groundwork for SURVEY.md 8f rank 3 (frame classification + BCH/Chase): the CPU restatement
oracle/ir_frame_oracle.c against the reference's own frame_decode() (frame_decode.c compiled unmodified
into oracle/_ref/libref_frame.so) on generated IRA / IBC frames -- clean, with correctable errors, with
errors only the Chase step can repair (incl. tied reliabilities), with too many errors, truncated, and on
random bits.  Every field of the flattened decoded_frame_t must agree, lat/lon included (same libm calls).
The device kernel of this row (k_classify_frames) is checked against this oracle in
tests/test_zz_gpu_classify.py; nothing in the product uses these files."""
import ctypes as C
import os

import numpy as np
import pytest

import importlib.util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("frame_gen", os.path.join(os.path.dirname(__file__), "frame_gen.py"))
fg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(fg)
ACCESS_DL, ACCESS_UL, make_ira, make_ibc, make_ida = fg.ACCESS_DL, fg.ACCESS_UL, fg.make_ira, fg.make_ibc, fg.make_ida
PORT_SO = os.path.join(ROOT, "oracle", "libir_frame_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_frame.so")

class Flat(C.Structure):
    _fields_ = [("ret", C.c_int32), ("type", C.c_int32), ("sat_id", C.c_int32), ("beam_id", C.c_int32),
                ("lat", C.c_double), ("lon", C.c_double), ("alt", C.c_int32), ("pos_xyz", C.c_int32 * 3),
                ("n_pages", C.c_int32), ("tmsi", C.c_uint32 * 12), ("msc_id", C.c_int32 * 12),
                ("timeslot", C.c_int32), ("sv_blocking", C.c_int32), ("bc_type", C.c_int32), ("iri_time", C.c_uint32)]


@pytest.fixture(scope="module")
def libs():
    from oracle import bindings as ob
    if not os.path.exists(PORT_SO):
        ob.build(port=True, ref=False)
    if not os.path.exists(REF_SO):
        if not os.path.exists("/root/reference/frame_decode.c"):
            pytest.skip("oracle/_ref/libref_frame.so not built and /root/reference absent")
        ob.build(port=False, ref=True)
    port, ref = C.CDLL(PORT_SO), C.CDLL(REF_SO)
    for f in (port.orc_frame_decode, ref.ref_frame_decode):
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Flat)]
    return port.orc_frame_decode, ref.ref_frame_decode


def _both(libs, bits, llr):
    bits = np.ascontiguousarray(bits, np.uint8)
    out = []
    for fn in libs:
        o = Flat()
        lp = None if llr is None else np.ascontiguousarray(llr, np.float32).ctypes.data_as(C.c_void_p)
        fn(bits.ctypes.data_as(C.c_void_p), lp, len(bits), C.byref(o))
        out.append(o)
    a, b = out
    assert bytes(a) == bytes(b), ({k: getattr(a, k) for k, _ in Flat._fields_ if not hasattr(getattr(a, k), "_length_")},
                                  {k: getattr(b, k) for k, _ in Flat._fields_ if not hasattr(getattr(b, k), "_length_")})
    return a


def test_clean_frames(libs):
    rng = np.random.default_rng(1)
    for n_pages in (0, 1, 3, 7):
        bits = make_ira(rng, n_pages)
        o = _both(libs, bits, None)
        assert o.ret == 1 and o.type == 1 and o.n_pages == n_pages
    for bc in (0, 1, 2, 3):
        for pairs in (1, 2, 3, 4):
            bits = make_ibc(rng, pairs, bc)
            o = _both(libs, bits, rng.uniform(0.1, 1.0, len(bits)))
            assert o.ret == 1 and o.type == 2 and o.bc_type == bc


def test_bit_errors_bch_and_chase(libs):
    rng = np.random.default_rng(2)
    decoded = 0
    for trial in range(400):
        bits = make_ira(rng, int(rng.integers(0, 5))) if trial % 2 else make_ibc(rng, int(rng.integers(1, 5)), int(rng.integers(0, 4)))
        bits = np.array(bits, np.uint8)
        llr = rng.uniform(0.2, 1.0, len(bits)).astype(np.float32)
        n_err = int(rng.integers(0, 14))
        pos = rng.choice(np.arange(24, len(bits)), n_err, replace=False)
        bits[pos] ^= 1
        llr[pos] = rng.uniform(0.0, 0.25, n_err)                 # errors are (mostly) the unreliable bits
        if trial % 5 == 0:
            llr = np.round(llr * 8) / 8                          # ties: the selection order must match exactly
        o = _both(libs, bits, llr if trial % 7 else None)
        decoded += o.ret
    assert decoded > 150                                         # the generator does produce decodable frames


def test_truncated_and_garbage(libs):
    rng = np.random.default_rng(3)
    full = make_ira(rng, 4)
    for n in (0, 10, 23, 24, 60, 119, 120, 121, 183, 184, 185, 250, len(full)):
        _both(libs, full[:n], None)
    full = make_ibc(rng, 4, 1)
    for n in (29, 30, 93, 94, 95, 157, 158, 159, 262 + 24, len(full)):
        _both(libs, full[:n], rng.uniform(0, 1, n))
    hits = 0
    for _ in range(300):
        n = int(rng.integers(24, 500))
        bits = rng.integers(0, 2, n).astype(np.uint8)
        if rng.integers(0, 2):
            bits[:24] = ACCESS_DL
        hits += _both(libs, bits, rng.uniform(0, 1, n)).ret
    assert hits < 30                                             # random payloads (almost) never pass three parity-checked blocks


# ================================================================= IDA bursts (ida_decode.c:543-662)
class FlatIda(C.Structure):
    _fields_ = [("ret", C.c_int32), ("ft", C.c_int32), ("lcw_ok", C.c_int32), ("lcw_ft", C.c_int32),
                ("lcw_code", C.c_int32), ("ec_lcw", C.c_int32), ("lcw3_val", C.c_uint32), ("da_ctr", C.c_int32),
                ("da_len", C.c_int32), ("cont", C.c_int32), ("payload_len", C.c_int32), ("crc_ok", C.c_int32),
                ("fixederrs", C.c_int32), ("bch_len", C.c_int32), ("stored_crc", C.c_uint16),
                ("computed_crc", C.c_uint16), ("payload", C.c_uint8 * 32), ("bch_stream", C.c_uint8 * 256)]


@pytest.fixture(scope="module")
def ida_libs(libs):
    port, ref = C.CDLL(PORT_SO), C.CDLL(REF_SO)
    for f in (port.orc_ida_decode, ref.ref_ida_decode):
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(FlatIda)]
    return port.orc_ida_decode, ref.ref_ida_decode


def _both_ida(ida_libs, bits, llr, direction=1):
    bits = np.ascontiguousarray(bits, np.uint8)
    out = []
    for fn in ida_libs:
        o = FlatIda()
        lp = None if llr is None else np.ascontiguousarray(llr, np.float32).ctypes.data_as(C.c_void_p)
        fn(bits.ctypes.data_as(C.c_void_p), lp, len(bits), direction, C.byref(o))
        out.append(o)
    a, b = out
    assert bytes(a) == bytes(b), ([(k, getattr(a, k), getattr(b, k)) for k, _ in FlatIda._fields_[:16] if getattr(a, k) != getattr(b, k)])
    return a


def test_ida_clean_and_crc(ida_libs):
    rng = np.random.default_rng(11)
    for da_len in (0, 1, 7, 20):
        for extra in (0, 4):
            o = _both_ida(ida_libs, make_ida(rng, da_len, extra_words=extra), None, direction=1 + da_len % 2)
            assert o.ret == 1 and o.ft == 2 and o.da_len == da_len and o.bch_len == 200 + 20 * extra
            if da_len > 0 and extra == 0:
                assert o.crc_ok == 1 and o.computed_crc == 0
    o = _both_ida(ida_libs, make_ida(rng, 12, good_crc=False), None)
    assert o.ret == 1 and o.crc_ok == 0
    assert _both_ida(ida_libs, make_ida(rng, 5, ft=3), None).ret == 0        # not an IDA frame type
    assert _both_ida(ida_libs, make_ida(rng, 5), None, direction=0).ret == 0  # direction undefined
    assert _both_ida(ida_libs, make_ida(rng, 21), None).ret == 0             # da_len > 20


def test_ida_bit_errors_and_truncation(ida_libs):
    rng = np.random.default_rng(12)
    ok = 0
    for trial in range(500):
        bits = np.array(make_ida(rng, int(rng.integers(0, 21)), extra_words=4 * int(rng.integers(0, 2))), np.uint8)
        llr = rng.uniform(0.2, 1.0, len(bits)).astype(np.float32)
        n_err = int(rng.integers(0, 16))
        pos = rng.choice(np.arange(24, len(bits)), n_err, replace=False)
        bits[pos] ^= 1
        llr[pos] = rng.uniform(0.0, 0.25, n_err)
        if trial % 4 == 0:
            llr = np.round(llr * 8) / 8
        n = len(bits)
        if trial % 6 == 0:
            n = 24 + 46 + 4 * int(rng.integers(0, (len(bits) - 70) // 4 + 1))   # cut on a 4-bit boundary (even symbol counts)
        ok += _both_ida(ida_libs, bits[:n], (llr[:n] if trial % 5 else None)).ret
    assert ok > 200
    for _ in range(200):                                                     # random bits: both must refuse alike
        n = 24 + 46 + 4 * int(rng.integers(31, 120))
        _both_ida(ida_libs, rng.integers(0, 2, n).astype(np.uint8), rng.uniform(0, 1, n))
