"""Generators of Iridium frames as demodulated bit arrays (test-side only): IRA and IBC frames
(access code, BCH(31,21)+parity blocks, two- and three-way interleaving: the inverse of frame_decode.c:156-199,
414-598) and IDA bursts (link control word, BCH(31,20) words in 124-bit blocks, CRC: the inverse of
ida_decode.c:193-394).  Used by the oracle, host-compiled and GPU classification tests."""
import numpy as np

ACCESS_DL = [0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 1, 1, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 1, 1]      # frame_decode.c:51-53
ACCESS_UL = [1, 1, 0, 0, 1, 1, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 1, 1, 1, 1, 1, 1, 0, 0]      # :54-56


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


LCW_FROM = [40, 39, 36, 35, 32, 31, 28, 27, 24, 23, 20, 19, 16, 15, 12, 11, 8, 7, 4, 3,
            41, 38, 37, 34, 33, 30, 29, 26, 25, 22, 21, 18, 17, 14, 13, 10, 9, 6, 5, 2,
            1, 46, 45, 44, 43, 42]                               # ida_decode.c:54-60


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


def make_ida(rng, da_len, ft=2, extra_words=0, good_crc=True, lcw_ft=None, lcw_code=None, lcw3=None):
    # link control word: three short BCH code words, permuted and dibit-swapped on the air
    l2d, l3d = int(rng.integers(0, 64)), int(rng.integers(0, 1 << 21))
    if lcw_ft is not None:
        l2d = (l2d & 0x0f) | (lcw_ft << 4)
    if lcw_code is not None:
        l2d = (l2d & 0x30) | lcw_code
    if lcw3 is not None:
        l3d = lcw3
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




def corpus(seed, n):
    """n (bits, llr or None, direction) cases: clean frames, correctable errors, Chase-only errors with tied
    reliabilities, too many errors, truncations and random bits -- the mix the oracle tests use"""
    rng = np.random.default_rng(seed)
    out = []
    for t in range(n):
        kind = t % 4
        if kind == 0:
            bits = make_ira(rng, int(rng.integers(0, 8)))
        elif kind == 1:
            bits = make_ibc(rng, int(rng.integers(1, 5)), int(rng.integers(0, 4)))
        elif kind == 2:
            bits = make_ida(rng, int(rng.integers(0, 22)), ft=2 if rng.integers(0, 8) else 3, extra_words=4 * int(rng.integers(0, 2)),
                            good_crc=bool(rng.integers(0, 4)))
        else:
            m = int(rng.integers(0, 250)) * 4
            bits = list(rng.integers(0, 2, m))
            if m >= 24 and rng.integers(0, 2):
                bits[:24] = ACCESS_DL
        bits = np.array(bits, np.uint8)
        llr = rng.uniform(0.2, 1.0, len(bits)).astype(np.float32)
        if len(bits) > 30:
            n_err = min(int(rng.integers(0, 16)) if t % 3 else 0, len(bits) - 24)
            pos = rng.choice(np.arange(24, len(bits)), n_err, replace=False)
            bits[pos] ^= 1
            llr[pos] = rng.uniform(0.0, 0.25, n_err)
        if t % 5 == 0:
            llr = np.round(llr * 8) / 8
        m = len(bits)
        if t % 6 == 5 and m > 70:
            m = 24 + 46 + 4 * int(rng.integers(0, (m - 70) // 4 + 1))
        out.append((bits[:m].copy(), None if t % 7 == 6 else llr[:m].copy(), int(rng.integers(0, 3)) if t % 9 == 8 else 1 + t % 2))
    return out


def planted_recordings(synth, seed=7):
    """Two short recordings whose bursts carry real frames: IRA / IBC frames on simplex channels (capture centred
    above 1626 MHz, where burst_downmix.c:764-766 allows up to 444 symbols) and IDA bursts on duplex channels
    (191 symbols at most, iridium.h:24 -- exactly one 382-bit IDA burst).  Returns [(recording, frames)]."""
    rng = np.random.default_rng(seed)
    simplex = [make_ira(rng, 3), make_ibc(rng, 4, 1), make_ira(rng, 0), make_ibc(rng, 3, 2)]
    duplex = [make_ida(rng, 9), make_ida(rng, 20), make_ida(rng, 0), make_ida(rng, 13, good_crc=False)]
    a = synth.make_recording(seed, duration_s=1.1, n_bursts=8, snr_db=(22.0, 28.0), frame_bits=simplex,
                             center_freq=1_632_000_000.0)
    b = synth.make_recording(seed + 1, duration_s=1.1, n_bursts=8, snr_db=(22.0, 28.0), frame_bits=duplex)
    return [(a, simplex), (b, duplex)]
