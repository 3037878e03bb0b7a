"""ctypes mirror of ir_frame_class_t (include/iridium_b200.h) and of the flattened structs of the frame oracle /
reference glue (oracle/ir_frame_oracle.c, oracle/ref_frame_glue.c), plus the field-for-field comparison the
classification tests share."""
import ctypes as C

import numpy as np


class FrameClass(C.Structure):
    _fields_ = [("frame_type", C.c_int32), ("sat_id", C.c_int32), ("beam_id", C.c_int32), ("lat", C.c_double),
                ("lon", C.c_double), ("alt", C.c_int32), ("pos_xyz", C.c_int32 * 3), ("n_pages", C.c_int32),
                ("tmsi", C.c_uint32 * 12), ("msc_id", C.c_int32 * 12), ("timeslot", C.c_int32),
                ("sv_blocking", C.c_int32), ("bc_type", C.c_int32), ("iri_time", C.c_uint32), ("ida_ok", C.c_int32),
                ("lcw_ft", C.c_int32), ("lcw_code", C.c_int32), ("ec_lcw", C.c_int32), ("lcw3_val", C.c_uint32),
                ("da_ctr", C.c_int32), ("da_len", C.c_int32), ("cont", C.c_int32), ("payload_len", C.c_int32),
                ("crc_ok", C.c_int32), ("fixederrs", C.c_int32), ("bch_len", C.c_int32), ("stored_crc", C.c_uint16),
                ("computed_crc", C.c_uint16), ("payload", C.c_uint8 * 32), ("bch_stream", C.c_uint8 * 256)]


class Flat(C.Structure):
    _fields_ = [("ret", C.c_int32), ("type", C.c_int32), ("sat_id", C.c_int32), ("beam_id", C.c_int32),
                ("lat", C.c_double), ("lon", C.c_double), ("alt", C.c_int32), ("pos_xyz", C.c_int32 * 3),
                ("n_pages", C.c_int32), ("tmsi", C.c_uint32 * 12), ("msc_id", C.c_int32 * 12),
                ("timeslot", C.c_int32), ("sv_blocking", C.c_int32), ("bc_type", C.c_int32), ("iri_time", C.c_uint32)]


class FlatIda(C.Structure):
    _fields_ = [("ret", C.c_int32), ("ft", C.c_int32), ("lcw_ok", C.c_int32), ("lcw_ft", C.c_int32),
                ("lcw_code", C.c_int32), ("ec_lcw", C.c_int32), ("lcw3_val", C.c_uint32), ("da_ctr", C.c_int32),
                ("da_len", C.c_int32), ("cont", C.c_int32), ("payload_len", C.c_int32), ("crc_ok", C.c_int32),
                ("fixederrs", C.c_int32), ("bch_len", C.c_int32), ("stored_crc", C.c_uint16),
                ("computed_crc", C.c_uint16), ("payload", C.c_uint8 * 32), ("bch_stream", C.c_uint8 * 256)]


def bind_checker(lib, prefix):
    """(bits, llr, direction) -> (Flat, FlatIda) from orc_* (oracle port) or ref_* (the reference itself)"""
    fdec, idec = getattr(lib, prefix + "frame_decode"), getattr(lib, prefix + "ida_decode")
    fdec.restype = idec.restype = C.c_int
    fdec.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Flat)]
    idec.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(FlatIda)]

    def run(bits, llr, direction):
        bits = np.ascontiguousarray(bits, np.uint8)
        lp = None if llr is None else np.ascontiguousarray(llr, np.float32).ctypes.data_as(C.c_void_p)
        a, b = Flat(), FlatIda()
        fdec(bits.ctypes.data_as(C.c_void_p), lp, len(bits), C.byref(a))
        idec(bits.ctypes.data_as(C.c_void_p), lp, len(bits), direction, C.byref(b))
        return a, b
    return run


_FRAME_FIELDS = ["sat_id", "beam_id", "lat", "lon", "alt", "n_pages", "timeslot", "sv_blocking", "bc_type", "iri_time"]
_IDA_FIELDS = ["lcw_ft", "lcw_code", "ec_lcw", "lcw3_val", "da_ctr", "da_len", "cont", "payload_len", "crc_ok",
               "fixederrs", "bch_len", "stored_crc", "computed_crc"]


def assert_same(got, flat, ida, where="", geo_tol=0.0):
    """got: FrameClass from the product's code; flat / ida: what frame_decode() / ida_decode() said.  geo_tol:
    absolute tolerance in degrees on lat / lon only (0 = bit-exact: host-compiled code shares the reference's C
    library; 1e-11 for the GPU, whose double atan2 is not the C library's) -- every other field is exact."""
    assert got.frame_type == (flat.type if flat.ret else 0), (where, "type", got.frame_type, flat.ret, flat.type)
    if flat.ret:
        for k in _FRAME_FIELDS:
            if k in ("lat", "lon") and geo_tol > 0:
                assert abs(getattr(got, k) - getattr(flat, k)) <= geo_tol, (where, k, getattr(got, k), getattr(flat, k))
                continue
            assert getattr(got, k) == getattr(flat, k), (where, k, getattr(got, k), getattr(flat, k))
        for k in ("pos_xyz", "tmsi", "msc_id"):
            assert list(getattr(got, k)) == list(getattr(flat, k)), (where, k)
    assert got.ida_ok == ida.ret, (where, "ida", got.ida_ok, ida.ret)
    if ida.ret:
        for k in _IDA_FIELDS:
            assert getattr(got, k) == getattr(ida, k), (where, k, getattr(got, k), getattr(ida, k))
        assert bytes(got.payload) == bytes(ida.payload), (where, "payload")
        assert bytes(got.bch_stream) == bytes(ida.bch_stream), (where, "bch_stream")
    else:
        # a refused frame leaves the IDA half untouched (zero)
        assert got.bch_len == 0 and got.payload_len == 0, where


# ---- the reference's own structs (frame_decode.h, ida_decode.h), for tests that drive its functions directly
class Lcw(C.Structure):                    # ida_decode.h:19-26
    _fields_ = [("ft", C.c_int), ("lcw_ok", C.c_int), ("lcw_ft", C.c_int), ("lcw_code", C.c_int),
                ("lcw3_val", C.c_uint32), ("ec_lcw", C.c_int)]


class IdaBurst(C.Structure):               # ida_decode.h:29-56
    _fields_ = [("timestamp", C.c_uint64), ("frequency", C.c_double), ("direction", C.c_int), ("magnitude", C.c_float),
                ("noise", C.c_float), ("level", C.c_float), ("confidence", C.c_int), ("n_symbols", C.c_int),
                ("da_ctr", C.c_int), ("da_len", C.c_int), ("cont", C.c_int), ("payload", C.c_uint8 * 32),
                ("payload_len", C.c_int), ("crc_ok", C.c_int), ("stored_crc", C.c_uint16), ("computed_crc", C.c_uint16),
                ("fixederrs", C.c_int), ("bch_stream", C.c_uint8 * 256), ("bch_len", C.c_int), ("lcw", Lcw),
                ("lcw_header", C.c_char * 128)]


class IdaSlot(C.Structure):                # ida_decode.h:59-67
    _fields_ = [("active", C.c_int), ("direction", C.c_int), ("frequency", C.c_double), ("last_timestamp", C.c_uint64),
                ("last_ctr", C.c_int), ("data", C.c_uint8 * 256), ("data_len", C.c_int)]


class IdaContext(C.Structure):             # ida_decode.h:72-74
    _fields_ = [("slots", IdaSlot * 16)]


IDA_CB = C.CFUNCTYPE(None, C.POINTER(C.c_uint8), C.c_int, C.c_uint64, C.c_double, C.c_int, C.c_float, C.c_void_p)
