"""Runs the GPU cases of the frame classification row (tests/gpu_classify_cases.py: k_classify_frames through the
C ABI against the oracle and the reference), of the linked drop-in program (tests/gpu_dropin_cases.py) and of the
time-block sharding of one stream (tests/gpu_block_cases.py) one per child process: a child process keeps a crash
(or a program the case started that still holds the GPU) from taking the whole pytest run with it, and the name sorts
this file last under `pytest -x`.  On a box with a CUDA device a case that SKIPS is a failure (its checker -- oracle/_ref,
the reference's libref_frame.so -- did not travel); the only skip allowed is the two-device case on a one-GPU box.
Every case's summary line is printed so that the driver's tail shows what ran."""
import os
import signal
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ["gpu_classify_cases.py::" + c for c in (
    "test_generated_frames_one_launch", "test_pipeline_classifies_planted_frames_from_device_memory",
    "test_parsed_output_of_a_run", "test_reference_named_entry_points", "test_classify_refuses_bad_arguments")]
CASES.append("gpu_dropin_cases.py::test_reference_main_linked_against_the_library")
CASES.append("gpu_dropin_cases.py::test_reference_detector_on_the_fft_plugin")   # -DUSE_GPU reference on gpu_burst_fft_*
CASES.append("gpu_block_cases.py::test_time_blocks_through_the_cuda_path")      # time blocks, merged
CASES.append("gpu_block_cases.py::test_one_process_driver_on_the_gpu")           # ir_multi_*: the same in one C call
CASES.append("gpu_block_cases.py::test_one_process_driver_parsed_on_the_gpu")    # ... with classification / --parsed text
CASES.append("gpu_block_cases.py::test_one_process_driver_independent_streams_on_the_gpu")   # config 5 in one process
CASES.append("gpu_block_cases.py::test_two_devices_blocks_and_streams")          # ir_multi_* owning two GPUs
CASES.append("gpu_fir_poison_cases.py::test_fir_ignores_what_shared_memory_held")  # NaN-filled shared memory under the FIR
CASES.append("gpu_crowded_boundary_cases.py::test_crowded_chunk_boundary")   # 112 bursts alive at a chunk boundary
MAY_SKIP = {"gpu_block_cases.py::test_two_devices_blocks_and_streams"}           # ... on a one-GPU box


@pytest.mark.parametrize("case", CASES)
def test_classification_on_the_gpu(case):
    # own session: on a timeout the whole group goes (the case, and any program it started that still holds the GPU)
    pr = subprocess.Popen([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
                           os.path.join("tests", case)], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                          start_new_session=True)
    try:
        out, err = pr.communicate(timeout=420)
    except subprocess.TimeoutExpired:
        os.killpg(pr.pid, signal.SIGKILL)
        out, err = pr.communicate()
        pytest.fail("timed out after 420 s:\n" + (out + err)[-4000:])
    r = subprocess.CompletedProcess(pr.args, pr.returncode, out, err)
    tail = (r.stdout + r.stderr)[-4000:]
    summary = [l for l in r.stdout.splitlines() if " passed" in l or " skipped" in l or " failed" in l or " error" in l]
    print(f"[{case}] {summary[-1] if summary else 'no summary line'}")
    assert r.returncode == 0, tail
    if "1 passed" not in r.stdout:
        assert case in MAY_SKIP and "skipped" in r.stdout, "the case did not run (a skip is a failure on a GPU box):\n" + tail
