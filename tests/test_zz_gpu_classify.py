"""Runs the GPU cases of the frame classification row (tests/gpu_classify_cases.py: k_classify_frames through the
C ABI against the oracle and the reference), of the linked drop-in program (tests/gpu_dropin_cases.py) and of the
time-block sharding of one stream (tests/gpu_block_cases.py) one per child process.  They were written after round 1's GPU minutes
were spent and had not run on a B200 when committed; a child process keeps a crash in not-yet-proven code from
taking the whole pytest run with it, and the name sorts this file last under `pytest -x`."""
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
CASES.append("gpu_block_cases.py::test_time_blocks_through_the_cuda_path")      # time blocks, merged
CASES.append("gpu_block_cases.py::test_one_process_driver_on_the_gpu")           # ir_multi_*: the same in one C call
CASES.append("gpu_block_cases.py::test_one_process_driver_parsed_on_the_gpu")    # ... with classification / --parsed text
CASES.append("gpu_block_cases.py::test_one_process_driver_independent_streams_on_the_gpu")   # config 5 in one process


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
    assert r.returncode == 0, tail
    assert "1 passed" in r.stdout or "skipped" in r.stdout, tail
