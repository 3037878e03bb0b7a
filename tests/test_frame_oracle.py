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


# ================================================================= IDA bursts (ida_decode.c:543-662)
class FlatIda(C.Structure):
    _fields_ = [("ret", C.c_int32), ("ft", C.c_int32), ("lcw_ok", C.c_int32), ("lcw_ft", C.c_int32),
                ("lcw_code", C.c_int32), ("ec_lcw", C.c_int32), ("lcw3_val", C.c_uint32), ("da_ctr", C.c_int32),
                ("da_len", C.c_int32), ("cont", C.c_int32), ("payload_len", C.c_int32), ("crc_ok", C.c_int32),
                ("fixederrs", C.c_int32), ("bch_len", C.c_int32), ("stored_crc", C.c_uint16),
                ("computed_crc", C.c_uint16), ("payload", C.c_uint8 * 32), ("bch_stream", C.c_uint8 * 256)]


LCW_FROM = [40, 39, 36, 35, 32, 31, 28, 27, 24, 23, 20, 19, 16, 15, 12, 11, 8, 7, 4, 3,
            41, 38, 37, 34, 33, 30, 29, 26, 25, 22, 21, 18, 17, 14, 13, 10, 9, 6, 5, 2,
            1, 46, 45, 44, 43, 42]                               # ida_decode.c:54-60


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


def _crc(data):
    crc = 0xFFFF
    for byte in data:
        crc ^= byte << 8
        for _ in range(8):
            crc = ((crc << 1) ^ 0x1021) & 0xFFFF if crc & 0x8000 else (crc << 1) & 0xFFFF
    return crc


def _cw20(d20):
    d = int("".join(map(str, d20)), 2)
    return [int(c) for c in format((d << 11) | _rem(3545, d << 11), "031b")]


def _interleave_n(h1, h2, ns):
    out = [0] * (2 * ns)
    for k, s in enumerate(range(ns - 1, 0, -2)):
        out[2 * s], out[2 * s + 1] = h1[2 * k], h1[2 * k + 1]
    for k, s in enumerate(range(ns - 2, -1, -2)):
        out[2 * s], out[2 * s + 1] = h2[2 * k], h2[2 * k + 1]
    return out


def make_ida(rng, da_len, ft=2, extra_words=0, good_crc=True):
    # link control word: three short BCH code words, permuted and dibit-swapped on the air
    l2d, l3d = int(rng.integers(0, 64)), int(rng.integers(0, 1 << 21))
    v1 = (ft << 4) | _rem(29, ft << 4)
    v2 = (l2d << 8) | _rem(465, l2d << 8)                      # 14 bits; only the top 13 are sent (the last is taken as 0)
    v3 = (l3d << 5) | _rem(41, l3d << 5)
    lcw_bits = _bits(v1, 7) + _bits(v2 >> 1, 13) + _bits(v3, 26)
    lcw = [0] * 46
    for i, src in enumerate(LCW_FROM):
        lcw[(src - 1) ^ 1] = lcw_bits[i]
    # payload stream: 20 header bits, 160 payload bits, CRC, padding to whole 20-bit words
    hdr = [int(b) for b in rng.integers(0, 2, 20)]
    hdr[3] = int(rng.integers(0, 2))
    hdr[5:8] = _bits(int(rng.integers(0, 8)), 3)
    hdr[11:16] = _bits(da_len, 5)
    hdr[17:20] = [0, 0, 0]
    pay = [int(b) for b in rng.integers(0, 2, 160)]
    head = hdr + [0] * 12 + pay
    crc = _crc(bytes(int("".join(map(str, head[i:i + 8])), 2) for i in range(0, 192, 8)))
    if not good_crc:
        crc ^= 0x0400
    stream = hdr + pay + _bits(crc, 16) + [int(b) for b in rng.integers(0, 2, 4 + 20 * extra_words)]
    words = [_cw20(stream[i:i + 20]) for i in range(0, len(stream), 20)]
    assert len(words) == 10 + extra_words
    body = []
    while len(words) >= 4:                                     # full blocks: the words are read 4th, 2nd, 3rd, 1st
        a, b, c, d = words[:4]
        words = words[4:]
        comb = d + b + c + a
        body += _interleave_n(comb[:62], comb[62:], 62)
    if words:                                                  # partial block of two words: halves swapped, first bits dropped
        assert len(words) == 2
        h2 = [int(rng.integers(0, 2))] + words[0]
        h1 = [int(rng.integers(0, 2))] + words[1]
        body += _interleave_n(h1, h2, 32)
    return list(ACCESS_DL) + lcw + body


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
