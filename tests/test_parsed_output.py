"""The `--parsed` sink (SURVEY.md 8f ranks 2+3): ir_format_lcw / ir_format_ida (host text in libiridium_b200.so)
against the reference's own ida_decode() + frame_output_print_ida() (frame_output.c, ida_decode.c compiled
unmodified into oracle/_ref/libref_frame.so, stdout captured), byte for byte -- every LCW type/code pair, every
payload length, good and bad CRCs, trailing bits, uplink and downlink, zero level -- and against golden lines the
reference produced (tests/golden/parsed_lines.json, made by tests/golden/make_golden_parsed.py), which keep the
check alive where oracle/_ref cannot be built."""
import ctypes as C
import importlib
import importlib.util
import json
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_frame.so")
GOLDEN = os.path.join(HERE, "golden", "parsed_lines.json")
BASE_NS = 1_700_000_000 * 1_000_000_000          # time origin of every line in these tests (a whole second)


def _load(name, sub=""):
    spec = importlib.util.spec_from_file_location(name, os.path.join(HERE, sub, name + ".py"))
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
    lib.fc_host_classify.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(fc.FrameClass)]
    return lib


@pytest.fixture(scope="module")
def product():
    pl = importlib.import_module("iridium-sniffer_b200.pipeline")
    L = pl.load_library()                       # host text functions: no GPU involved
    L.ir_format_ida.restype = C.c_int
    L.ir_format_ida.argtypes = [C.c_char_p, C.c_size_t, C.c_uint64, C.POINTER(pl.Frame), C.c_void_p]
    L.ir_format_lcw.restype = C.c_int
    L.ir_format_lcw.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p]
    return pl, L


def product_line(shim, product, bits, llr, meta):
    """bits -> class (the kernel's arithmetic, host-compiled) -> the library's IDA line and LCW header"""
    pl, L = product
    bits = np.ascontiguousarray(bits, np.uint8)
    cls = fc.FrameClass()
    lp = None if llr is None else np.ascontiguousarray(llr, np.float32).ctypes.data_as(C.c_void_p)
    shim.fc_host_classify(bits.ctypes.data_as(C.c_void_p), lp, len(bits), meta["direction"], C.byref(cls))
    if not cls.ida_ok:
        return "", ""
    f = pl.Frame()
    f.timestamp, f.center_frequency, f.direction = meta["timestamp"], meta["center_frequency"], meta["direction"]
    f.magnitude, f.noise, f.level, f.confidence = meta["magnitude"], meta["noise"], meta["level"], meta["confidence"]
    f.n_payload_symbols, f.n_symbols, f.n_bits = meta["n_payload_symbols"], meta["n_payload_symbols"] + 12, len(bits)
    buf, hdr = C.create_string_buffer(2048), C.create_string_buffer(256)
    n = L.ir_format_ida(buf, len(buf), BASE_NS, C.byref(f), C.byref(cls))
    m = L.ir_format_lcw(hdr, len(hdr), C.byref(cls))
    assert n > 0 and m == 111
    assert L.ir_format_ida(buf, n, BASE_NS, C.byref(f), C.byref(cls)) == -1      # one byte short: refused, no overrun
    L.ir_format_ida(buf, len(buf), BASE_NS, C.byref(f), C.byref(cls))
    return buf.value.decode(), hdr.value.decode()


def cases(seed=21):
    """(bits, llr, meta): every LCW (type, code) pair and a spread of everything else the line shows.  All are
    382-bit bursts (bch_len 200), the longest a duplex channel carries (iridium.h:24); past bch_len 256 the
    reference prints beyond its own bch_stream array, which nothing can be asked to reproduce."""
    rng = np.random.default_rng(seed)
    out = []
    k = 0
    for lcw_ft in range(4):
        for lcw_code in range(16):
            for rep in range(3):
                da_len = [0, 20, int(rng.integers(1, 20))][rep]
                bits = np.array(fg.make_ida(rng, da_len, good_crc=bool(rng.integers(0, 3)),
                                            lcw_ft=lcw_ft, lcw_code=lcw_code,
                                            lcw3=[0, (1 << 21) - 1, None][rep]), np.uint8)
                llr = rng.uniform(0.2, 1.0, len(bits)).astype(np.float32)
                if k % 4 == 0:                                   # a few correctable errors: fixed-up words print the same text
                    pos = rng.choice(np.arange(70, len(bits)), 3, replace=False)
                    bits[pos] ^= 1
                    llr[pos] = 0.01
                meta = dict(direction=1 + k % 2, timestamp=BASE_NS + int(rng.integers(0, 3_000_000_000)),
                            center_frequency=float(rng.uniform(1.616e9, 1.6265e9)), magnitude=float(np.float32(rng.uniform(5, 60))),
                            noise=float(np.float32(rng.uniform(-130, -90))), level=float(np.float32(0.0 if k % 17 == 5 else rng.uniform(1e-3, 2.0))),
                            confidence=int(rng.integers(0, 101)), n_payload_symbols=int(rng.integers(-1, 180)))
                out.append((bits, llr if k % 3 else None, meta))
                k += 1
    return out


def test_ida_lines_equal_the_references(shim, product):
    if not os.path.exists(REF_SO):
        if not os.path.exists("/root/reference/frame_output.c"):
            pytest.skip("oracle/_ref/libref_frame.so not built and /root/reference absent (golden lines still checked)")
        from oracle import bindings as ob
        ob.build(port=False, ref=True)
    ref = C.CDLL(REF_SO)
    ref.ref_print_prime.argtypes = [C.c_uint64]
    ref.ref_print_ida.restype = C.c_int
    ref.ref_print_ida.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_double, C.c_float, C.c_float,
                                  C.c_float, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_char_p]
    ref.ref_print_prime(BASE_NS + 5)
    n_lines = 0
    for bits, llr, meta in cases():
        out, hdr = C.create_string_buffer(4096), C.create_string_buffer(128)
        lp = None if llr is None else llr.ctypes.data_as(C.c_void_p)
        n = ref.ref_print_ida(bits.ctypes.data_as(C.c_void_p), lp, len(bits), meta["direction"], meta["timestamp"],
                              meta["center_frequency"], meta["magnitude"], meta["noise"], meta["level"], meta["confidence"],
                              meta["n_payload_symbols"], out, len(out), hdr)
        assert n >= 0
        line, lcw = product_line(shim, product, bits, llr, meta)
        assert line == out.value.decode()
        if n:
            assert lcw == hdr.value.decode()
            n_lines += 1
    assert n_lines > 150


def test_ida_lines_equal_golden(shim, product):
    with open(GOLDEN) as fh:
        gold = json.load(fh)
    assert len(gold) >= 100
    for g in gold:
        bits = np.array([int(c) for c in g["bits"]], np.uint8)
        line, lcw = product_line(shim, product, bits, None, g["meta"])
        assert line == g["line"] and lcw == g["lcw_header"]


def test_parsed_run_equals_the_reference_program(shim, product, port, synth, tmp_path):
    """End to end on the CPU: the reference program with --parsed over a recording whose bursts carry IDA frames,
    against oracle path (detect -> downmix -> demod) -> the kernel's host-compiled classification -> the library's
    line formatter.  Everything after the time stamp must be identical (the program stamps with CLOCK_REALTIME,
    and its four downmix threads may swap neighbouring lines, hence the sort)."""
    from oracle import bindings as ob
    if not os.path.exists(ob.REF_BIN):
        if not os.path.exists("/root/reference/main.c"):
            pytest.skip("oracle/_ref/iridium-sniffer not built and /root/reference absent")
        ob.build(port=False, ref=True)
    rec, _ = fg.planted_recordings(synth)[1]
    path = str(tmp_path / "ida.cf32")
    rec.iq.tofile(path)
    r = subprocess.run([ob.REF_BIN, "-f", path, "--format=cf32", "-r", str(rec.sample_rate), "-c", str(int(rec.center_freq)),
                        "--parsed"], capture_output=True, text=True, check=True)
    want = sorted(l.split(" ", 3)[3] for l in r.stdout.splitlines())
    assert sum(l.startswith("IDA:") for l in r.stdout.splitlines()) >= 6
    res, _ = port.run(rec.iq, center_frequency=rec.center_freq, sample_rate=rec.sample_rate, start_time_ns=BASE_NS)
    got = []
    for fr in res:
        meta = {k: fr[k] for k in ("direction", "timestamp", "center_frequency", "magnitude", "noise", "level", "confidence",
                                   "n_payload_symbols")}
        line, _ = product_line(shim, product, fr["bits"], None, meta)
        if not line:                                   # not an IDA frame: --parsed prints its RAW line
            line = port.format_raw("x", BASE_NS, fr) + "\n"
        got.append(line.rstrip("\n").split(" ", 3)[3])
    assert sorted(got) == want


def test_formatter_survives_records_no_classifier_would_produce(product):
    """garbage in a class record (negative or huge lengths, codes out of range) must not crash or overrun"""
    pl, L = product
    f = pl.Frame()
    f.timestamp = BASE_NS + 1
    buf = C.create_string_buffer(4096)
    for bch_len, da_len, lcw_ft, lcw_code in ((-5, 3, 0, 0), (100000, 20, 7, 99), (19, 0, -1, -1), (256, 31, 2, 3), (200, 100, 1, 1)):
        c = fc.FrameClass()
        c.ida_ok, c.bch_len, c.da_len, c.lcw_ft, c.lcw_code, c.lcw3_val = 1, bch_len, da_len, lcw_ft, lcw_code, 0xFFFFFFFF
        n = L.ir_format_ida(buf, len(buf), BASE_NS, C.byref(f), C.byref(c))
        assert 0 < n < 1000 and buf.value.decode().startswith("IDA: p-1700000000 ") and buf.value.endswith(b"\n")
    c = fc.FrameClass()
    assert L.ir_format_ida(buf, len(buf), BASE_NS, C.byref(f), C.byref(c)) == -1      # not an IDA frame
    assert L.ir_format_lcw(buf, len(buf), C.byref(c)) == -1
