"""A few seconds of GPU: k_classify_frames through ir_classify_frames on generated IRA / IBC / IDA frames against the
oracle port (and the reference's functions when oracle/_ref travelled), results appended line by line to
gpurun_out/quick_classify.log so that a cut-off run still leaves what it saw.  No torch, no scipy."""
import ctypes as C
import importlib
import importlib.util
import os
import sys
import time

T0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "quick_classify.log"), "a")


def say(*a):
    line = "[%6.2fs] " % (time.time() - T0) + " ".join(str(x) for x in a)
    print(line, flush=True)
    LOG.write(line + "\n")
    LOG.flush()
    os.fsync(LOG.fileno())


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tests", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


say("start")
fg, fc = _load("frame_gen"), _load("frame_class_types")
pl = importlib.import_module("iridium-sniffer_b200.pipeline")
L = pl.load_library()
say("library loaded, devices:", L.ir_device_count())
checkers = [("port", fc.bind_checker(C.CDLL(os.path.join(ROOT, "oracle", "libir_frame_oracle.so")), "orc_"))]
ref_so = os.path.join(ROOT, "oracle", "_ref", "libref_frame.so")
if os.path.exists(ref_so):
    checkers.append(("reference", fc.bind_checker(C.CDLL(ref_so), "ref_")))
for n in (int(sys.argv[1]) if len(sys.argv) > 1 else 400, 3000):
    cases = fg.corpus(202, n)
    say("corpus", n)
    for label, group in (("llr", [c for c in cases if c[1] is not None]), ("hard", [c for c in cases if c[1] is None])):
        t = time.time()
        got = pl.classify_frames(group)
        dt = time.time() - t
        bad = dec = 0
        first = None
        for o, (bits, llr, direction) in zip(got, group):
            o = fc.FrameClass.from_buffer_copy(bytes(o))
            dec += (o.frame_type != 0) + o.ida_ok
            for name, chk in checkers:
                try:
                    fc.assert_same(o, *chk(bits, llr, direction), where=name, geo_tol=1e-11)
                except AssertionError as e:
                    bad += 1
                    first = first or repr(e)[:300]
        say("classify", label, "frames", len(group), "decoded", dec, "mismatches", bad, "call %.1f ms" % (dt * 1e3),
            "checkers", [c[0] for c in checkers], first or "")
say("done")
