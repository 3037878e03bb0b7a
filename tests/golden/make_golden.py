"""Regenerate tests/golden/*.json from the reference itself (oracle/_ref).

Run in the build container (needs /root/reference to have been compiled by
`make -C oracle ref`):  python tests/golden/make_golden.py
The IQ is not stored -- only generator seeds and the reference's answers.
"""
import importlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as ob  # noqa: E402

synth = importlib.import_module("iridium-sniffer_b200.synth")
HERE = os.path.dirname(os.path.abspath(__file__))


def lines_for(rec, extra=None):
    with tempfile.NamedTemporaryFile(suffix="." + rec.fmt, dir="/tmp") as f:
        (rec.iq.view(np.float32) if rec.fmt == "cf32" else rec.iq).tofile(f.name)
        lines, err, _ = ob.run_ref_binary(f.name, rec.fmt, rec.sample_rate, rec.center_freq, extra)
    tagged = [l for l in err.splitlines() if "tagged" in l]
    return [ob.parse_raw(l) for l in lines], tagged


def main():
    out = {}
    rec = synth.make_recording(1234, duration_s=1.5, n_bursts=12)
    res, tagged = lines_for(rec)
    out["config1_cf32_10MHz_seed1234"] = dict(
        gen=dict(seed=1234, sample_rate=10_000_000, duration_s=1.5, n_bursts=12, fmt="cf32"),
        tagged=tagged, lines=res, truth_bits=[b.bits for b in rec.truth])
    res_ng, _ = lines_for(rec, ["--no-gardner"])
    out["config1_no_gardner"] = dict(lines=res_ng)
    rec3 = synth.make_recording(3, sample_rate=12_000_000, duration_s=1.2, n_bursts=8, fmt="ci16",
                                center_freq=1_621_000_000.0)
    res3, tagged3 = lines_for(rec3)
    out["config3_ci16_12MHz_seed3"] = dict(
        gen=dict(seed=3, sample_rate=12_000_000, duration_s=1.2, n_bursts=8, fmt="ci16",
                 center_freq=1_621_000_000.0),
        tagged=tagged3, lines=res3, truth_bits=[b.bits for b in rec3.truth])
    with open(os.path.join(HERE, "ref_lines.json"), "w") as f:
        json.dump(out, f, indent=1)
    for k, v in out.items():
        print(k, len(v["lines"]), v.get("tagged"))

    # stage vectors for two bursts of config 1 (reference's own stage outputs)
    ref = ob.Ref()
    bursts = ref.detect(rec.iq)
    vec = {}
    for b in bursts[:2]:
        fr = ref.downmix(b)
        dm = ref.demod(fr)
        vec[f"burst{b['id']}_hdr"] = np.array([b["id"], b["start"], b["stop"], b["center_bin"]], np.int64)
        vec[f"burst{b['id']}_frame"] = fr["samples"]
        vec[f"burst{b['id']}_bits"] = dm["bits"]
        vec[f"burst{b['id']}_scalars"] = np.array(
            [fr["center_frequency"], dm["center_frequency"], dm["level"], dm["confidence"],
             b["magnitude"], b["noise"], fr["uw_start"]], np.float64)
    np.savez_compressed(os.path.join(HERE, "config1_stage_vectors.npz"), **vec)


if __name__ == "__main__":
    main()
