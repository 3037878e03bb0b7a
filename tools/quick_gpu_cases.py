"""The GPU cases of tests/gpu_classify_cases.py, tests/gpu_dropin_cases.py and tests/gpu_block_cases.py in ONE process, most valuable first,
each result appended to gpurun_out/quick_gpu_cases.log as soon as it is known (for a GPU call of a few seconds).
    python tools/quick_gpu_cases.py"""
import ctypes as C
import importlib
import importlib.util
import os
import sys
import tempfile
import time
import traceback

T0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "quick_gpu_cases.log"), "a")


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
gc, gd = _load("gpu_classify_cases"), _load("gpu_dropin_cases")
synth = importlib.import_module("iridium-sniffer_b200.synth")
pl = importlib.import_module("iridium-sniffer_b200.pipeline")
fc = gc.fc
checkers = [("port", fc.bind_checker(C.CDLL(gc.PORT_SO), "orc_"))]
if os.path.exists(gc.REF_SO):
    checkers.append(("reference", fc.bind_checker(C.CDLL(gc.REF_SO), "ref_")))
planted = gc.fg.planted_recordings(synth)
gc.fg.planted_recordings = lambda s, seed=7: planted
gd.fg.planted_recordings = lambda s, seed=7: planted
say("recordings made; checkers", [c[0] for c in checkers])


def step(name, fn):
    t = time.time()
    try:
        fn()
        say("PASS", name, "%.2f s" % (time.time() - t))
    except BaseException as e:  # noqa: BLE001 -- log whatever it was, go on
        say("FAIL", name, "%.2f s" % (time.time() - t), repr(e)[:400])
        LOG.write(traceback.format_exc()[-3000:] + "\n")
        LOG.flush()


step("pipeline_classifies_planted_frames_from_device_memory",
     lambda: gc.test_pipeline_classifies_planted_frames_from_device_memory(pl, checkers, synth))
step("parsed_output_of_a_run", lambda: gc.test_parsed_output_of_a_run(pl, checkers, synth))
step("reference_named_entry_points", lambda: gc.test_reference_named_entry_points(pl))
step("classify_refuses_bad_arguments", lambda: gc.test_classify_refuses_bad_arguments(pl))


def timing():
    rec, _ = planted[1]
    p = pl.Pipeline(sample_rate=rec.sample_rate, center_frequency=rec.center_freq)
    p.run_host(rec.iq, rec.fmt)
    ms = []
    for _ in range(5):
        n = len(p.classify())
        ms.append(p.classify_ms())
    say("k_classify_frames device ms per launch (", n, "frames ):", ["%.4f" % m for m in ms])
    p.close()


step("classify_kernel_timing", timing)


def dropin(i, extra):
    rec, _ = planted[i]
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "rec.cf32")
        rec.iq.tofile(path)
        a, b = gd._run(gd.REF_BIN, path, rec, extra), gd._run(gd.NEW_BIN, path, rec, extra)
        n = gd._compare(a, b)
        say("   dropin", i, extra, "lines", n, "IDA", sum(l.startswith("IDA:") for l in b))


step("dropin_duplex_parsed", lambda: dropin(1, ["--parsed"]))
step("dropin_simplex_raw", lambda: dropin(0, []))
step("dropin_duplex_raw", lambda: dropin(1, []))
step("dropin_simplex_parsed", lambda: dropin(0, ["--parsed"]))
step("generated_frames_one_launch", lambda: gc.test_generated_frames_one_launch(pl, checkers))
gb = _load("gpu_block_cases")              # one stream in time blocks, ir_multi_*
step("time_blocks_through_the_cuda_path", gb.test_time_blocks_through_the_cuda_path)
step("one_process_driver", gb.test_one_process_driver_on_the_gpu)
step("one_process_driver_parsed", gb.test_one_process_driver_parsed_on_the_gpu)
step("one_process_driver_independent_streams", gb.test_one_process_driver_independent_streams_on_the_gpu)
say("done")
