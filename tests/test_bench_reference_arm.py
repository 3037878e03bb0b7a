"""bench.py --impl reference runs the reference's own program (oracle/_ref) on a bounded sample and
prints ONE JSON line with the contract's keys -- no GPU involved, so it is checked here."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    from oracle import bindings as ob
    if not os.path.exists(ob.REF_BIN):
        if not os.path.exists("/root/reference/burst_detect.c"):
            pytest.skip("oracle/_ref not built and /root/reference absent")
        ob.build(port=False, ref=True)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--ref-seconds", "1.2",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "IQ Msamples/s detect->RAW" and d["unit"] == "Msamples/s"
    assert d["higher_is_better"] is True and d["value"] > 1.0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["config"]["raw_lines"] >= 1
