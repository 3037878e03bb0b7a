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
