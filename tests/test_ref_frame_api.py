"""The reference-named interface of the frame row (include/ir_ref_api.h: frame_decode.h / ida_decode.h), CPU side:
  * ir_fill_decoded_frame / ir_fill_ida_burst turn a class record into the reference's structs -- compared BYTE FOR
    BYTE with what the reference's own frame_decode() / ida_decode() leave in decoded_frame_t / ida_burst_t
    (lcw_header text, padding and all); the class records come from the kernel's arithmetic compiled for the host;
  * ida_reassemble / ida_reassemble_flush against the reference's, on random burst traffic: same callbacks, same
    return values, same context bytes after every call;
  * the public bit helpers of frame_decode.h.
frame_decode() / ida_decode() themselves run their frame through the GPU: tests/test_zz_gpu_classify.py."""
import ctypes as C
import importlib
import importlib.util
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_frame.so")


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(HERE, name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


fg = _load("frame_gen")
fc = _load("frame_class_types")


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_SO) or not hasattr(C.CDLL(REF_SO), "ref_ida_decode_raw"):
        if not os.path.exists("/root/reference/frame_decode.c"):
            pytest.skip("oracle/_ref/libref_frame.so not built and /root/reference absent")
        from oracle import bindings as ob
        ob.build(port=False, ref=True)
    return C.CDLL(REF_SO)


@pytest.fixture(scope="module")
def lib():
    return importlib.import_module("iridium-sniffer_b200.pipeline").load_library()


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("fc") / "libfc_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-x", "c++",
                    os.path.join(HERE, "fc_host_shim.cpp"), "-o", out], check=True)
    s = C.CDLL(out)
    s.fc_host_classify.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(fc.FrameClass)]
    return s


def test_struct_layouts(ref):
    assert ref.ref_sizeof_ida_burst() == C.sizeof(fc.IdaBurst)
    assert ref.ref_sizeof_ida_context() == C.sizeof(fc.IdaContext)


def test_filled_structs_equal_the_references_byte_for_byte(ref, lib, shim):
    pl = importlib.import_module("iridium-sniffer_b200.pipeline")
    rng = np.random.default_rng(31)
    nd, nb = ref.ref_sizeof_decoded_frame(), ref.ref_sizeof_ida_burst()
    raw_args = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_double, C.c_float, C.c_float, C.c_float, C.c_int,
                C.c_int, C.c_void_p]
    ref.ref_frame_decode_raw.argtypes = raw_args
    ref.ref_ida_decode_raw.argtypes = raw_args
    lib.ir_fill_decoded_frame.argtypes = [C.POINTER(pl.Frame), C.c_void_p, C.c_void_p]
    lib.ir_fill_ida_burst.argtypes = [C.POINTER(pl.Frame), C.c_void_p, C.c_void_p]
    seen = {"ira": 0, "ibc": 0, "ida": 0, "none": 0}
    for bits, llr, direction in fg.corpus(55, 800):
        ts, freq = int(rng.integers(1, 2**62)), float(rng.uniform(1.616e9, 1.6265e9))
        mag, noise, level = (float(np.float32(v)) for v in (rng.uniform(5, 60), rng.uniform(-130, -90), rng.uniform(0, 2)))
        conf, npay = int(rng.integers(0, 101)), len(bits) // 2 - 12
        lp = None if llr is None else llr.ctypes.data_as(C.c_void_p)
        want_d, want_b = C.create_string_buffer(b"\xaa" * nd, nd), C.create_string_buffer(b"\xaa" * nb, nb)
        r1 = ref.ref_frame_decode_raw(bits.ctypes.data_as(C.c_void_p), lp, len(bits), direction, ts, freq, mag, noise, level, conf, npay, want_d)
        r2 = ref.ref_ida_decode_raw(bits.ctypes.data_as(C.c_void_p), lp, len(bits), direction, ts, freq, mag, noise, level, conf, npay, want_b)
        cls = fc.FrameClass()
        shim.fc_host_classify(bits.ctypes.data_as(C.c_void_p), lp, len(bits), direction, C.byref(cls))
        f = pl.Frame()
        f.timestamp, f.center_frequency, f.direction, f.magnitude, f.noise, f.level = ts, freq, direction, mag, noise, level
        f.confidence, f.n_payload_symbols, f.n_symbols, f.n_bits = conf, npay, npay + 12, len(bits)
        got_d, got_b = C.create_string_buffer(b"\x55" * nd, nd), C.create_string_buffer(b"\x55" * nb, nb)
        lib.ir_fill_decoded_frame(C.byref(f), C.byref(cls), got_d)
        r2m = lib.ir_fill_ida_burst(C.byref(f), C.byref(cls), got_b)
        assert (cls.frame_type != 0) == bool(r1) and r2m == r2
        assert got_d.raw == want_d.raw
        assert got_b.raw == want_b.raw
        seen["ira" if cls.frame_type == 1 else "ibc" if cls.frame_type == 2 else "ida" if r2 else "none"] += 1
    assert min(seen.values()) > 50, seen


def _traffic(rng, n):
    """bursts on a handful of channels: message starts, continuations, gaps, wrong counters, bad CRCs; every so
    often a storm of message starts on forty channels at once (all 16 slots taken: evictions) or one long message
    (more fragments than the 256 bytes of a slot hold)"""
    chans = [(1 + int(rng.integers(0, 2)), 1.620e9 + 41666.0 * k) for k in range(40)]
    state = {}
    t = 10**18
    out = []

    def burst(ch, ctr, da_len, cont, crc_ok=1, jitter=100.0):
        b = fc.IdaBurst()
        b.timestamp = t
        b.frequency = ch[1] + float(rng.uniform(-jitter, jitter))
        b.direction = ch[0]
        b.magnitude = float(rng.uniform(5, 40))
        b.da_ctr, b.da_len, b.cont, b.crc_ok = ctr, da_len, cont, crc_ok
        for i in range(da_len):
            b.payload[i] = int(rng.integers(0, 256))
        b.payload_len = da_len or 20
        return b

    while len(out) < n:
        mode = rng.random()
        if mode < 0.02:                                         # storm: many messages open at once
            for k in rng.permutation(40)[:int(rng.integers(17, 40))]:
                t += int(rng.integers(1, 4)) * 1_000_000
                out.append(burst(chans[int(k)], 0, 20, 1))
                state[chans[int(k)]] = 1
            continue
        if mode < 0.04:                                         # one long message, 90 ms apart
            ch = chans[int(rng.integers(0, 40))]
            for k in range(int(rng.integers(10, 18))):
                t += 90_000_000
                out.append(burst(ch, k % 8, 20, 1))
            t += 90_000_000
            out.append(burst(ch, (k + 1) % 8, 5, 0))
            state[ch] = 0
            continue
        t += int(rng.choice([2_000_000, 10_000_000, 20_000_000, 40_000_000, 300_000_000]))
        ch = chans[int(rng.integers(0, 4))] if rng.random() < 0.8 else chans[int(rng.integers(0, len(chans)))]
        ctr = state.get(ch, 0)
        if rng.random() < 0.15:
            ctr = int(rng.integers(0, 8))                       # out of sequence
        b = burst(ch, ctr, int(rng.choice([0, 1, 7, 20, 20, 20])), int(rng.random() < 0.85), int(rng.random() < 0.9),
                  300.0 if rng.random() < 0.2 else 100.0)
        if rng.random() < 0.03:
            b.timestamp -= int(rng.integers(1, 10**6))          # now and then out of order
        state[ch] = (ctr + 1) % 8 if b.cont else 0
        out.append(b)
    return out


def test_reassembly_equals_the_references(ref, lib):
    rng = np.random.default_rng(41)
    for fn in (ref.ida_reassemble, lib.ida_reassemble):
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(fc.IdaContext), C.POINTER(fc.IdaBurst), fc.IDA_CB, C.c_void_p]
    for fn in (ref.ida_reassemble_flush, lib.ida_reassemble_flush):
        fn.restype = None
        fn.argtypes = [C.POINTER(fc.IdaContext), C.c_uint64]
    logs = ([], [])

    def mk(log):
        return fc.IDA_CB(lambda data, n, ts, fr, d, mag, user: log.append((bytes(data[:n]), n, ts, fr, d, mag)))
    cbs = (mk(logs[0]), mk(logs[1]))
    ctxs = (fc.IdaContext(), fc.IdaContext())
    emitted = full_house = 0
    for i, b in enumerate(_traffic(rng, 6000)):
        r0 = ref.ida_reassemble(C.byref(ctxs[0]), C.byref(b), cbs[0], None)
        r1 = lib.ida_reassemble(C.byref(ctxs[1]), C.byref(b), cbs[1], None)
        assert r0 == r1, i
        if i % 3 == 0:                                          # main.c:354 flushes after every frame; any cadence must agree
            ref.ida_reassemble_flush(C.byref(ctxs[0]), b.timestamp)
            lib.ida_reassemble_flush(C.byref(ctxs[1]), b.timestamp)
        assert bytes(ctxs[0]) == bytes(ctxs[1]), i
        emitted += r0
        full_house += all(s.active for s in ctxs[0].slots)       # the next new message evicts the oldest
    assert logs[0] == logs[1]
    multi = sum(1 for l in logs[0] if l[1] > 20)
    longest = max(l[1] for l in logs[0])
    assert emitted > 100 and multi > 30 and longest > 236 and full_house > 0, (emitted, multi, longest, full_house)


def test_bit_helpers(ref, lib):
    rng = np.random.default_rng(43)
    for L in (ref, lib):
        L.gf2_remainder.restype = C.c_uint32
        L.gf2_remainder.argtypes = [C.c_uint32, C.c_uint32]
        L.bits_to_uint.restype = C.c_uint32
        L.bits_to_uint.argtypes = [C.c_void_p, C.c_int]
        L.uint_to_bits.argtypes = [C.c_uint32, C.c_void_p, C.c_int]
        L.bch_31_21_correct.restype = C.c_int
        L.bch_31_21_correct.argtypes = [C.c_uint32, C.POINTER(C.c_uint32)]
    ref.frame_decode_init()
    lib.frame_decode_init()
    for poly in (29, 41, 465, 1207, 3545):
        for v in [0, 1, poly, poly - 1, 2**31 - 1] + [int(x) for x in rng.integers(0, 2**31, 200)]:
            assert ref.gf2_remainder(poly, v) == lib.gf2_remainder(poly, v)
    for s in range(1024):
        a, b = C.c_uint32(123), C.c_uint32(123)
        assert ref.bch_31_21_correct(s, C.byref(a)) == lib.bch_31_21_correct(s, C.byref(b)) and a.value == b.value
    for n in (1, 7, 21, 32):
        bits = rng.integers(0, 2, n).astype(np.uint8)
        v = ref.bits_to_uint(bits.ctypes.data_as(C.c_void_p), n)
        assert v == lib.bits_to_uint(bits.ctypes.data_as(C.c_void_p), n)
        o1, o2 = np.full(n, 9, np.uint8), np.full(n, 9, np.uint8)
        ref.uint_to_bits(v, o1.ctypes.data_as(C.c_void_p), n)
        lib.uint_to_bits(v, o2.ctypes.data_as(C.c_void_p), n)
        assert o1.tolist() == o2.tolist() == bits.tolist()


_SINK_SCRIPT = r'''
import ctypes as C, importlib, importlib.util, os, sys, numpy as np
ROOT, REF_SO, out_dir = sys.argv[1], sys.argv[2], sys.argv[3]
sys.path.insert(0, ROOT)
def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tests", name + ".py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m); return m
fg, fc = _load("frame_gen"), _load("frame_class_types")
from oracle import bindings as ob
ref = C.CDLL(REF_SO)
lib = importlib.import_module("iridium-sniffer_b200.pipeline").load_library()
raw_args = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_double, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p]
ref.ref_ida_decode_raw.argtypes = raw_args
libc = C.CDLL(None)
FI = None if sys.argv[4] == "auto" else C.c_char_p(sys.argv[4].encode())      # borrowed by both: must stay alive
rng = np.random.default_rng(3)
items = []
t = 1_723_456_789_123_456_789
for k in range(60):
    t += int(rng.integers(1, 10**8))
    da = int(rng.integers(0, 21))
    bits = np.array(fg.make_ida(rng, da, good_crc=bool(k % 3)), np.uint8) if k % 2 else rng.integers(0, 2, int(rng.integers(0, 400))).astype(np.uint8)
    f = ob.RefDemodFrame()
    f.id, f.timestamp, f.center_frequency, f.direction = int(rng.integers(0, 10**10)), t, float(rng.uniform(1.616e9, 1.6265e9)), 1 + k % 2
    f.magnitude, f.noise, f.level, f.confidence = float(rng.uniform(5, 60)), float(rng.uniform(-130, -90)), float(rng.uniform(0, 2)), int(rng.integers(0, 101))
    f.n_payload_symbols, f.n_symbols, f.n_bits = len(bits) // 2 - 12, len(bits) // 2, len(bits)
    f.bits = bits.ctypes.data_as(C.POINTER(C.c_uint8))
    burst = fc.IdaBurst()
    ok = ref.ref_ida_decode_raw(bits.ctypes.data_as(C.c_void_p), None, len(bits), f.direction, f.timestamp, f.center_frequency,
                                f.magnitude, f.noise, f.level, f.confidence, f.n_payload_symbols, C.byref(burst))
    items.append((f, bits, burst if ok else None))
for name, L in (("ref", ref), ("lib", lib)):
    L.frame_output_init.argtypes = [C.c_char_p]
    L.frame_output_print.argtypes = [C.POINTER(ob.RefDemodFrame)]
    L.frame_output_print_ida.argtypes = [C.POINTER(fc.IdaBurst)]
    fd = os.open(os.path.join(out_dir, name + ".txt"), os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
    libc.fflush(None)
    saved = os.dup(1)
    os.dup2(fd, 1)
    L.frame_output_init(FI)
    for f, bits, burst in items:
        if burst is not None:
            L.frame_output_print_ida(C.byref(burst))       # main.c:328-331 with --parsed
        else:
            L.frame_output_print(C.byref(f))
    libc.fflush(None)
    os.dup2(saved, 1)
    os.close(fd)
'''


@pytest.mark.parametrize("file_info", ["auto", "rec-17"])
def test_per_line_sinks_equal_the_references(ref, tmp_path, file_info):
    """frame_output_init / frame_output_print / frame_output_print_ida under the reference's names: the bytes on
    stdout for a mixed run of RAW and IDA lines, time origin and automatic file_info included (in a child
    process: both implementations fix their time origin at the first line they ever print)."""
    import sys
    r = subprocess.run([sys.executable, "-c", _SINK_SCRIPT, ROOT, REF_SO, str(tmp_path), file_info], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    a, b = open(tmp_path / "ref.txt", "rb").read(), open(tmp_path / "lib.txt", "rb").read()
    assert a == b and a.count(b"\n") == 60 and a.count(b"IDA: ") >= 20 and a.count(b"RAW: ") >= 20
    assert (b"RAW: rec-17 " in a) == (file_info != "auto") and (b"RAW: i-1723456789-t1 " in a) == (file_info == "auto")
