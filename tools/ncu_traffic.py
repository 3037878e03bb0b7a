"""profiles/r2_kernel_traffic.json out of an ncu pass over ONE device-resident run of the bench recording
(tools/dev_timeline.py 60 1 under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
--clock-control none --csv`): DRAM bytes per unit of work for every kernel group bench.py reports, so that
`roofline.traffic` is a reading of this build and not a literal.

    python tools/ncu_traffic.py gpurun_out/r2_traffic_launches.csv gpurun_out/r2_traffic_run.json profiles/r2_kernel_traffic.json
The second file is the JSON line dev_timeline.py prints last (frames / tiles / bursts of the run)."""
import collections
import csv
import json
import sys

GROUPS = {   # bench.py's kernel groups -> (name fragments, unit key of the run, unit name)
    "k_detect_fft": (("k_detect_fft",), "det_frames", "detector frame"),
    "k_detect_scan": (("k_seg_", "k_detect_classify", "k_detect_scan"), "det_frames", "detector frame"),
    "k_fir": (("k_fir",), "tiles", "FIR tile"),
    "k_chain": (("k_chain",), "bursts", "burst"),
    "k_demod": (("k_demod",), "bursts", "burst"),
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3,
         "nsecond": 1e-9, "second": 1.0}


def main(launches, run_json, out):
    run = json.load(open(run_json))
    rows = list(csv.reader(l for l in open(launches) if l.startswith('"')))
    h = rows[0]
    ki, mi, ui, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value")
    acc = collections.defaultdict(lambda: {"bytes": 0.0, "s": 0.0, "launches": 0})
    for r in rows[1:]:
        name = r[ki]
        grp = next((g for g, (frags, _, _) in GROUPS.items() if any(f in name for f in frags)), None)
        if grp is None:
            continue
        v = float(r[vi].replace(",", "")) * SCALE.get(r[ui], 1.0)
        if r[mi].startswith("dram__bytes"):
            acc[grp]["bytes"] += v
        elif r[mi].startswith("gpu__time_duration"):
            acc[grp]["s"] += v
            acc[grp]["launches"] += 1
    res = {}
    for g, a in acc.items():
        units = run[GROUPS[g][1]]
        res[g] = {"dram_bytes_per_unit": a["bytes"] / max(units, 1), "unit": GROUPS[g][2], "units_in_capture": units,
                  "dram_bytes_in_capture": a["bytes"], "kernel_seconds_in_capture_serialised": a["s"], "launches": a["launches"],
                  "source": "profiles/" + launches.split("/")[-1]}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:4])
