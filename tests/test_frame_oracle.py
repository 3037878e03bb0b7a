"""Groundwork for SURVEY.md 8f rank 3 (frame classification + BCH/Chase): the CPU restatement
oracle/ir_frame_oracle.c against the reference's own frame_decode() (frame_decode.c compiled unmodified
into oracle/_ref/libref_frame.so) on generated IRA / IBC frames -- clean, with correctable errors, with
errors only the Chase step can repair (incl. tied reliabilities), with too many errors, truncated, and on
random bits.  Every field of the flattened decoded_frame_t must agree, lat/lon included (same libm calls).
No device path exists for this row yet; nothing in the product uses these files."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PORT_SO = os.path.join(ROOT, "oracle", "libir_frame_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_frame.so")

ACCESS_DL = [0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 1, 1, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 1, 1]      # frame_decode.c:51-53
ACCESS_UL = [1, 1, 0, 0, 1, 1, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 1, 1, 1, 1, 1, 1, 0, 0]      # :54-56


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


# ---------------------------------------------------------------- a frame generator (test-side only)
def _rem(poly, v):
    deg = poly.bit_length() - 1
    while v >> deg:
        v ^= poly << (v.bit_length() - 1 - deg)
    return v


def _block(data21):
    """21 data bits -> 32-bit block: systematic BCH(31,21) with generator 1207, then overall even parity"""
    d = int("".join(map(str, data21)), 2)
    code = (d << 10) | _rem(1207, d << 10)
    bits = [int(c) for c in format(code, "031b")]
    return bits + [sum(bits) & 1]


def _interleave2(b1, b2):
    out = [0] * 64
    for k in range(16):
        for blk, s in ((b1, 31 - 2 * k), (b2, 30 - 2 * k)):
            out[2 * s], out[2 * s + 1] = blk[2 * k], blk[2 * k + 1]
    return out


def _interleave3(b1, b2, b3):
    out = [0] * 96
    for k in range(16):
        for blk, s in ((b1, 47 - 3 * k), (b2, 46 - 3 * k), (b3, 45 - 3 * k)):
            out[2 * s], out[2 * s + 1] = blk[2 * k], blk[2 * k + 1]
    return out


def _bits(v, n):
    return [int(c) for c in format(v & ((1 << n) - 1), "0%db" % n)]


def make_ira(rng, n_pages):
    x, y, z = (int(v) for v in rng.integers(-2048, 2048, 3))
    s12 = lambda v: [1 if v < 0 else 0] + _bits(v + 2048 if v < 0 else v, 11)
    stream = _bits(int(rng.integers(0, 128)), 7) + _bits(int(rng.integers(0, 64)), 6) + s12(x) + s12(y) + s12(z)
    stream += [int(b) for b in rng.integers(0, 2, 63 - len(stream))]
    for _ in range(n_pages):
        stream += _bits(int(rng.integers(0, 2**32)), 32) + [int(b) for b in rng.integers(0, 2, 10)]
    stream += [1] * 42                                           # terminator page
    blocks = [_block(stream[i:i + 21]) for i in range(0, len(stream), 21)]
    assert len(blocks) % 2 == 1
    body = _interleave3(*blocks[:3])
    for i in range(3, len(blocks), 2):
        body += _interleave2(blocks[i], blocks[i + 1])
    return list(ACCESS_DL) + body


def make_ibc(rng, n_pairs, bc_type):
    hv = {0: 0, 1: 29, 2: 39, 3: 58}[bc_type]                    # the 6-bit multiples of the BCH(7,3) generator 29
    stream = _bits(int(rng.integers(0, 128)), 7) + _bits(int(rng.integers(0, 64)), 6) + [int(b) for b in rng.integers(0, 2, 29)]
    if n_pairs > 1:
        stream += _bits(int(rng.integers(0, 3)), 6) + [int(b) for b in rng.integers(0, 2, 36)]
    stream += [int(b) for b in rng.integers(0, 2, 42 * max(0, n_pairs - 2))]
    blocks = [_block(stream[i:i + 21]) for i in range(0, 42 * n_pairs, 21)]
    body = _bits(hv, 6)
    for i in range(0, len(blocks), 2):
        body += _interleave2(blocks[i], blocks[i + 1])
    return list(ACCESS_DL if rng.integers(0, 2) else ACCESS_UL) + body


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
