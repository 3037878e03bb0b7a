"""Host-side logic of libiridium_b200.so that needs no GPU: the RAW: line formatter against the
reference binary's own lines (tests/golden/ref_lines.json, frame_output.c:160-199), and the chunk
planner of ir_pipeline_run_*."""
import ctypes as C
import importlib
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "ref_lines.json")


def _lib():
    pl = importlib.import_module("iridium-sniffer_b200.pipeline")
    if not os.path.exists(pl.LIB_PATH):
        pl.build_library()
    return pl, pl.load_library()


def test_format_raw_reproduces_reference_lines():
    pl, L = _lib()
    gold = json.load(open(GOLD))
    n = 0
    for cfg in gold.values():
        for g in cfg["lines"]:
            # the reference's own format string (frame_output.c:176-190) over the fields parsed from its line
            want = "RAW: %s %012.4f %010d N:%05.2f%+06.2f I:%011d %3d%% %.5f %3d %s\n" % (
                g["file_info"], g["ts_ms"], g["freq_hz"], g["magnitude"], g["noise"], g["id"], g["confidence"],
                g["level"], g["n_payload"], g["bits"])
            f = pl.Frame()
            f.id = g["id"]
            f.timestamp = int(round(g["ts_ms"] * 1e6))
            f.center_frequency = float(g["freq_hz"])
            f.magnitude, f.noise = g["magnitude"], g["noise"]
            f.confidence = g["confidence"]
            f.level = g["level"]
            f.n_payload_symbols = g["n_payload"]
            f.n_symbols = g["n_payload"] + 12
            f.n_bits = len(g["bits"])
            b = np.frombuffer(g["bits"].encode(), np.uint8) - ord("0")
            buf = C.create_string_buffer(4096)
            k = L.ir_format_raw(buf, 4096, g["file_info"].encode(), 0, C.byref(f), b.ctypes.data_as(C.c_void_p))
            assert k == len(want) and buf.value.decode() == want, (buf.value.decode(), want)
            n += 1
    assert n >= 20


def _plan(L, n, chunk, N):
    ends = (C.c_size_t * 4096)()
    k = L.ir_plan_chunks(n, chunk, N, ends, 4096)
    assert k >= 0
    return [int(ends[i]) for i in range(k)]


def test_chunk_plan_halves_the_last_chunk():
    _, L = _lib()
    N, Mi = 8192, 1 << 20
    e = _plan(L, 600_000_000, 16 * Mi, N)
    sizes = np.diff([0] + e)
    assert e[-1] == 600_000_000 and all(s > 0 for s in sizes)
    assert all(x % N == 0 for x in e[:-1])                     # every piece but the last: whole frames
    assert all(s == 16 * Mi for s in sizes[:35])                # 35 full chunks, then 12.8 Mi in halves
    tail = sizes[35:]
    assert len(tail) >= 3 and tail[-1] <= 2 * Mi and all(tail[i] + 2 * N >= tail[i + 1] for i in range(len(tail) - 1))


def test_chunk_plan_edges():
    _, L = _lib()
    N, Mi = 8192, 1 << 20
    assert _plan(L, 0, 16 * Mi, N) == []
    assert _plan(L, 5000, 16 * Mi, N) == [5000]                # shorter than one frame: one ragged piece
    assert _plan(L, Mi, Mi, N) == [Mi]                         # chunk == minimum piece: no halving
    e = _plan(L, 3 * Mi + 12345, Mi, N)
    assert e == [Mi, 2 * Mi, 3 * Mi, 3 * Mi + 12345]
    e = _plan(L, 40_000_000, 1000, 16384)                      # chunk below one frame is rounded up to one frame
    assert e[0] == 16384 and e[-1] == 40_000_000


def test_fixed_point_formatter_is_printf():
    """ir_format_fixed (what the RAW: line is built with) == printf's %f on the line's four conversions, over random
    values, values next to rounding boundaries, exact ties (dyadic fractions: ties go to even) and the tiny / huge
    ranges where it hands over to printf"""
    _, L = _lib()
    rng = np.random.default_rng(7)
    buf = C.create_string_buffer(512)
    cases = []
    for dec, width, zp, plus, fmt in ((4, 12, 1, 0, "%012.4f"), (2, 5, 1, 0, "%05.2f"), (2, 6, 1, 1, "%+06.2f"), (5, 0, 0, 0, "%.5f")):
        xs = list(rng.uniform(-200.0, 200.0, 4000)) + list(rng.uniform(0, 6.0e7, 4000)) + list(rng.uniform(-1e-4, 1e-4, 2000))
        xs += list(np.float32(rng.uniform(-150, 60, 4000)).astype(np.float64))                     # float fields, promoted
        xs += [k / 2.0 ** j for k in range(-40, 41) for j in range(0, 9)]                           # exact ties and friends
        step = 10.0 ** -dec
        for k in range(-300, 300, 7):
            for e in (-1, 0, 1):
                xs.append(np.nextafter(k * step + step / 2, np.inf if e > 0 else -np.inf) if e else k * step + step / 2)
        xs += [0.0, -0.0, 1e15, -1e15, 9.007199254740991e15, 1e16, 1e300, 5e-324, 0.99999999, 9.999995, 99.995, 0.125, 0.375, 2.5, 3.5]
        for x in xs:
            cases.append((float(x), dec, width, zp, plus, fmt))
    bad = 0
    for x, dec, width, zp, plus, fmt in cases:
        k = L.ir_format_fixed(buf, x, dec, width, zp, plus)
        got = buf.value.decode()
        want = fmt % x
        assert k == len(got)
        if got != want:
            bad += 1
            assert bad < 1, (x, fmt, got, want)
    assert len(cases) > 50000
