#!/usr/bin/env python
"""bench.py -- IQ Msamples/s through detect -> downmix -> DQPSK demod -> RAW lines.

Contract (see the task statement / DESIGN.md section 6):
  python bench.py --gpus N --steps K --warmup W          our CUDA path
  python bench.py --impl reference ...                   the reference's own CPU path (oracle/_ref)
One "step" = one pass of the whole path over one synthetic 10 MHz cf32 recording
(BASELINE.json configs[1]; every rank of an N-GPU run gets its own recording = configs[4]).
Rank 0 prints ONE JSON line.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "IQ Msamples/s detect->RAW"
UNIT = "Msamples/s"
FS = 10_000_000


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_recording_gpu(torch, synth, seed, duration_s, bursts_per_s, device):
    """Config-2 style recording generated on the GPU (noise) + CPU-made burst waveforms."""
    n = int(duration_s * FS)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    iq = torch.empty((n, 2), dtype=torch.float32, device=device)
    step = 1 << 26
    for o in range(0, n, step):
        m = min(step, n - o)
        iq[o:o + m].normal_(0.0, 0.01, generator=g)
    rng = np.random.default_rng(seed)
    up = FS // 250_000
    pool = [synth.burst_waveform(rng, up, 179, False) for _ in range(48)]
    pool_dev = [(torch.from_numpy(np.stack([w.real, w.imag], 1).astype(np.float32)).to(device), b) for w, b in pool]
    raster = 1e6 / 24.0
    kmax = int((FS / 2 - 200e3) / raster)
    chans = [k * raster for k in range(-kmax, kmax + 1) if abs(k * raster) >= 60e3]
    lead = 512 * 8192 / FS + 0.02
    blen = len(pool[0][0])
    nb = int(bursts_per_s * (duration_s - lead))
    starts = np.sort(rng.uniform(lead, duration_s - blen / FS - 0.03, nb))
    last = {}
    truth = []
    kk = torch.arange(blen, device=device, dtype=torch.float64)
    for t0 in starts:
        for _ in range(32):
            ch = chans[int(rng.integers(0, len(chans)))]
            if t0 - last.get(ch, -1.0) >= 0.027 + blen / FS:
                break
        else:
            continue
        last[ch] = float(t0)
        w, bits = pool_dev[int(rng.integers(0, len(pool_dev)))]
        snr = float(rng.uniform(12.0, 25.0))
        amp = 0.01 * 10.0 ** (snr / 20.0)
        f = ch + float(rng.uniform(-3e3, 3e3))
        ph = float(rng.uniform(0, 2 * np.pi))
        s0 = int(round(t0 * FS))
        arg = (2 * np.pi * f / FS) * kk + ph
        c, s = torch.cos(arg).float(), torch.sin(arg).float()
        seg = iq[s0:s0 + blen]
        seg[:, 0] += amp * (w[:, 0] * c - w[:, 1] * s)
        seg[:, 1] += amp * (w[:, 0] * s + w[:, 1] * c)
        truth.append(bits)
    return iq, truth


def reduce_over_ranks(torch, dist, device, times, counts):
    """Job-level numbers from per-rank ones: MAX over ranks for every time, SUM for every count.
    dist is None for a single process.  (tests/test_multirank_gloo.py runs this on gloo.)"""
    t = torch.tensor(list(times), dtype=torch.float64, device=device)
    c = torch.tensor(list(counts), dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()], [int(x) for x in c.tolist()]


def shard_seed(world, rank):
    """Each rank decodes its own independent recording (different seed -> different bursts)."""
    return 2 + 8 * (world > 1) + rank


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    synth = importlib.import_module("iridium-sniffer_b200.synth")
    pl = importlib.import_module("iridium-sniffer_b200.pipeline")

    t_gen = time.time()
    iq_dev, truth = make_recording_gpu(torch, synth, shard_seed(world, rank), args.seconds, args.bursts_per_s, dev)
    n = iq_dev.shape[0]
    host = torch.empty((n, 2), dtype=torch.float32, pin_memory=True)
    host.copy_(iq_dev)
    torch.cuda.synchronize()
    t_gen = time.time() - t_gen
    truth_set = set(truth)

    p = pl.Pipeline(sample_rate=FS, device=local, start_time_ns=1_700_000_000_000_000_000)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident leg (value)
    # The timed loops call the C ABI only (run + per-step stats struct); result conversion to
    # Python objects happens once, after the clocks stop.
    iq_ptr, host_ptr = iq_dev.data_ptr(), host.data_ptr()
    for _ in range(args.warmup):
        p.run_device_raw(iq_ptr, n, "cf32")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms, launches, stage = 0.0, 0, {}
    for _ in range(args.steps):
        p.run_device_raw(iq_ptr, n, "cf32")
        st = p.stats()
        dev_ms += st["ms_total"]
        launches += st["kernel_launches"]
        for k in ("ms_detect_fft", "ms_detect_scan", "ms_downmix_fir", "ms_downmix_chain", "ms_demod"):
            stage[k] = stage.get(k, 0.0) + st[k]
    barrier()
    wall = time.perf_counter() - t0
    res = p.results()
    # ---------------- end-to-end leg: pinned host IQ -> RAW text lines
    for _ in range(args.warmup):
        p.run_host_raw(host_ptr, n, "cf32")
        p.raw_text_len("b200")
    barrier()
    t1 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(args.steps):
        p.run_host_raw(host_ptr, n, "cf32")
        p.raw_text_len("b200")                  # every RAW: line of the step, in the library's text buffer
        st = p.stats()
        h2d += st["h2d_bytes"]; d2h += st["d2h_bytes"]
    barrier()
    wall_e2e = time.perf_counter() - t1
    text = bytes(p.raw_text_view())
    n_lines = text.count(b"\n")
    assert text.startswith(b"RAW: b200 ") and n_lines == len(p.results().frames)
    clocks = sampler.stop() if rank == 0 else None
    # context for e2e: the bare pinned->device copy of one step's input (PCIe floor of this box)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    floor_ms = []
    for _ in range(3):
        torch.cuda.synchronize()
        ev0.record()
        iq_dev.copy_(host, non_blocking=True)
        ev1.record()
        torch.cuda.synchronize()
        floor_ms.append(ev0.elapsed_time(ev1))
    h2d_floor_ms = min(floor_ms)

    (wall, wall_e2e, dev_s), (n_bursts, n_frames, launches) = reduce_over_ranks(
        torch, dist if world > 1 else None, dev, [wall, wall_e2e, dev_ms / 1e3],
        [len(res.bursts), len(res.frames), launches])

    ok_bits = sum("".join(map(str, f["bits"])) in truth_set for f in res.frames)
    if rank == 0:
        peaks, peak_kind = measured_peaks()
        total = n * world * args.steps
        value = total / wall / 1e6
        K = args.steps
        # per-kernel view: algorithmic bytes over summed launch time (rank 0)
        sum_n = sum(b["num_samples"] for b in res.bursts if b["dec_len"] >= 100)
        sum_dec = sum(b["dec_len"] for b in res.bursts if b["dec_len"] >= 100)
        sum_fl = sum(b["frame_len"] for b in res.bursts if b["downmix_status"] == 0)
        kern = {
            "k_detect_fft": (12.0 * n, stage["ms_detect_fft"] / K),
            "k_detect_scan": (4.0 * n, stage["ms_detect_scan"] / K),      # classify + snapshot + k_detect_scan_stream
            "k_fir": (8.0 * sum_n + 8.0 * sum_dec, stage["ms_downmix_fir"] / K),
            "k_chain": (8.0 * (2 * sum_dec + sum_fl), stage["ms_downmix_chain"] / K),
            "k_demod": (8.0 * sum_fl + 5.0 * sum(f["n_bits"] for f in res.frames), stage["ms_demod"] / K),
        }
        # DRAM traffic per launch from the committed ncu capture (profiles/r1b_prof_stream_summary.csv,
        # dram__bytes_read.sum + dram__bytes_write.sum); the state-machine figure is for a 1536-frame launch
        ncu_traffic = {"k_detect_fft": (134.3e6 + 37.6e6, "2048-frame launch"),
                       "k_detect_scan": (19.7e6 + 50.4e6 + 0.2e6, "1536-frame launch: k_detect_scan_stream 19.7 MB + k_detect_classify 50.6 MB"),
                       "k_fir": (103.9e6 + 7.8e6, "3428-tile launch"), "k_chain": (7.0e6, "126-burst launch"),
                       "k_demod": (2.2e6, "126-burst launch")}
        dom = max(kern, key=lambda k: kern[k][1])
        ach = kern[dom][0] / (kern[dom][1] * 1e-3) / 1e9 if kern[dom][1] > 0 else 0.0
        roof = {"bound": "hbm", "kernel": dom, "achieved": round(ach, 2), "peak": peaks["hbm_gbs"],
                "peak_kind": peak_kind + (" burst copy bandwidth" if peak_kind == "measured" else ""),
                "unit": "GB/s", "frac": round(ach / peaks["hbm_gbs"], 5), "traffic": ncu_traffic[dom][0],
                "traffic_note": ncu_traffic[dom][1] + " (ncu, profiles/r1b_prof_stream_summary.csv)",
                "bound_note": ("the state machine is a serial recurrence over frames: one warp walks bitmaps "
                               "(1/16 of the magnitude bytes) at ~0.5 us per frame; latency-bound, not a bandwidth kernel"
                               if dom == "k_detect_scan" else ""),
                "ms_per_launch_sum": round(kern[dom][1], 4),
                "whole_path": {"alg_bytes": res.stats["alg_bytes"],
                               "achieved": round(res.stats["alg_bytes"] / (dev_s / K) / 1e9, 2),
                               "frac": round(res.stats["alg_bytes"] / (dev_s / K) / 1e9 / peaks["hbm_gbs"], 5)},
                "kernels": {k: {"alg_gb": round(v[0] / 1e9, 4), "ms": round(v[1], 4),
                                "gbs": round(v[0] / (v[1] * 1e-3) / 1e9, 1) if v[1] > 0 else None}
                            for k, v in kern.items()}}
        cpu = cpu_baseline_port(host.numpy().view(np.complex64).reshape(-1), args.cpu_seconds)
        out = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": args.warmup, "ms_per_step": round(wall / K * 1e3, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"synthetic 10 MHz cf32 recording, {args.seconds:g} s ({n} samples) per GPU, "
                                   f"8192-pt detect FFT, ~{args.bursts_per_s:g} bursts/s, full path on "
                                   f"{world}xB200 (one independent stream per GPU, no NCCL on the data path)",
                       "samples_per_gpu": n, "bursts": n_bursts, "raw_frames": n_frames,
                       "bits_matching_ground_truth": f"{ok_bits}/{len(res.frames)} (rank 0)",
                       "l2": "inputs larger than L2 (no flush needed)",
                       "device_ms_per_step": round(dev_s / K * 1e3, 3), "gen_s": round(t_gen, 1)},
            "e2e": {"value": round(total / wall_e2e / 1e6, 2), "unit": UNIT,
                    "h2d_bytes_per_step": h2d // K, "d2h_bytes_per_step": d2h // K,
                    "ms_per_step": round(wall_e2e / K * 1e3, 3), "raw_lines_per_step": n_lines,
                    "h2d_copy_alone_ms": round(h2d_floor_ms, 3),
                    "h2d_copy_alone_gbs": round(n * 8 / h2d_floor_ms / 1e6, 1),
                    "api": "ir_pipeline_run_host (pinned host IQ -> frames) + ir_pipeline_format_raw_all"},
            "bursts_per_s": round(n_bursts / (wall / K), 1),
            "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(out))
    p.close()
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline_port(iq: np.ndarray, seconds: float):
    """Oracle restatement (single thread) on the first `seconds` of the same recording."""
    from oracle import bindings as ob
    port = ob.Port()
    m = min(iq.shape[0], int(seconds * FS))
    t = time.perf_counter()
    res, st = port.run(iq[:m], sample_rate=FS)
    dt = time.perf_counter() - t
    return {"value": round(m / dt / 1e6, 2), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"first {m / FS:g} s ({m} samples) of the rank-0 recording: "
                      f"{st['n_bursts']} bursts, {st['n_results']} frames, "
                      f"detect {st['t_detect_s']:.2f}s downmix {st['t_downmix_s']:.2f}s demod {st['t_demod_s']:.2f}s"}


def run_reference(args):
    """The reference's own program (oracle/_ref/iridium-sniffer: unmodified sources + FFT shim,
    AVX2 kernels, its fixed 1+4+1 thread graph) on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import bindings as ob
    synth = importlib.import_module("iridium-sniffer_b200.synth")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not os.path.exists(ob.REF_BIN):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/iridium-sniffer not built"}))
        return
    secs = args.ref_seconds
    rec = synth.make_recording(2, sample_rate=FS, duration_s=secs, n_bursts=int(args.bursts_per_s * max(secs - 0.45, 0.1)),
                               waveform_pool=32)
    path = "/dev/shm/ir_bench_ref.cf32" if os.path.isdir("/dev/shm") else "/tmp/ir_bench_ref.cf32"
    rec.iq.view(np.float32).tofile(path)
    n = rec.n_samples

    def one():
        t = time.perf_counter()
        # stdout/stderr go to files: a pipe nobody drains would block the child after 64 KiB
        fo, fe = open(path + ".out", "wb"), open(path + ".err", "wb")
        pr = subprocess.Popen([ob.REF_BIN, "-f", path, "--format=cf32", "-r", str(FS), "--file-info=ref"],
                              stdout=fo, stderr=fe)
        busy = {}
        deadline = time.time() + 600
        while pr.poll() is None:          # per-thread CPU time: robust to the 1 s exit quantum (main.c:405,794)
            try:
                for tid in os.listdir(f"/proc/{pr.pid}/task"):
                    f = open(f"/proc/{pr.pid}/task/{tid}/stat").read().rsplit(")", 1)[1].split()
                    busy[tid] = (int(f[11]) + int(f[12])) / os.sysconf("SC_CLK_TCK")
            except Exception:
                pass
            time.sleep(0.02)
            if time.time() > deadline:
                pr.kill()
                break
        pr.wait()
        wall = time.perf_counter() - t
        fo.close(); fe.close()
        out = open(path + ".out", "rb").read().decode(errors="replace")
        os.remove(path + ".out"); os.remove(path + ".err")
        lines = [l for l in out.splitlines() if l.startswith("RAW:")]
        return wall, (max(busy.values()) if busy else wall), len(lines), len(busy)

    for _ in range(args.warmup if args.warmup < 2 else 1):
        one()
    walls, busys, nl, nth = [], [], 0, 0
    for _ in range(args.steps):
        w, b, nl, nth = one()
        walls.append(w); busys.append(b)
    os.remove(path)
    cores = os.cpu_count()
    bottleneck = float(np.mean(busys))
    value = n / bottleneck / 1e6
    out = {
        "impl": "reference", "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(bottleneck * 1e3, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"synthetic 10 MHz cf32 recording, bounded sample of {secs:g} s ({n} samples), "
                               "8192-pt detect FFT, reference CPU path (AVX2 kernels, FFT shim in place of FFTW)",
                   "timing": "samples / CPU time of the busiest thread (the detector): the reference's pipeline "
                             "throughput without its 1 s exit quantum; wall-clock figure in cpu_baseline.sample",
                   "raw_lines": nl},
        "cpu_baseline": {"value": round(value, 2), "unit": UNIT, "cores": min(7, cores), "kind": "reference",
                         "sample": f"{secs:g} s of signal; threads seen {nth}; host cores {cores}; "
                                   f"wall {np.mean(walls):.2f} s -> {n / np.mean(walls) / 1e6:.1f} Msps wall; "
                                   "libfftw3f absent -> oracle/shim FFT"},
        "e2e": {"value": round(value, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seconds", type=float, default=60.0, help="signal seconds per GPU per step")
    ap.add_argument("--bursts-per-s", type=float, default=100.0)
    ap.add_argument("--cpu-seconds", type=float, default=8.0, help="signal seconds given to the CPU oracle")
    ap.add_argument("--ref-seconds", type=float, default=10.0, help="signal seconds per reference-arm step")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
