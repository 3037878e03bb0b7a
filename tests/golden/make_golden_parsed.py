"""Writes tests/golden/parsed_lines.json: IDA frames (bits + frame metadata) and the exact line the reference's
ida_decode() + frame_output_print_ida() print for them (oracle/_ref/libref_frame.so = the reference's sources
compiled unmodified).  Run here, where /root/reference exists:  python tests/golden/make_golden_parsed.py"""
import ctypes as C
import importlib.util
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("test_parsed_output", os.path.join(os.path.dirname(HERE), "test_parsed_output.py"))
tp = importlib.util.module_from_spec(spec)
spec.loader.exec_module(tp)

from oracle import bindings as ob  # noqa: E402

ob.build(port=False, ref=True)
ref = C.CDLL(tp.REF_SO)
ref.ref_print_prime.argtypes = [C.c_uint64]
ref.ref_print_ida.restype = C.c_int
ref.ref_print_ida.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_double, C.c_float, C.c_float,
                              C.c_float, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_char_p]
ref.ref_print_prime(tp.BASE_NS + 5)
gold = []
for bits, llr, meta in tp.cases(seed=77):
    out, hdr = C.create_string_buffer(4096), C.create_string_buffer(128)
    n = ref.ref_print_ida(bits.ctypes.data_as(C.c_void_p), None, len(bits), meta["direction"], meta["timestamp"],
                          meta["center_frequency"], meta["magnitude"], meta["noise"], meta["level"], meta["confidence"],
                          meta["n_payload_symbols"], out, len(out), hdr)
    if n > 0:
        gold.append(dict(bits="".join(map(str, bits)), meta=meta, line=out.value.decode(), lcw_header=hdr.value.decode()))
with open(os.path.join(HERE, "parsed_lines.json"), "w") as fh:
    json.dump(gold, fh, indent=0)
print(len(gold), "lines")
