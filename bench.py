#!/usr/bin/env python
"""bench.py -- IQ Msamples/s through detect -> downmix -> DQPSK demod -> RAW lines.

Contract (see the task statement / DESIGN.md section 6):
  python bench.py --gpus N --steps K --warmup W          our CUDA path
  python bench.py --impl reference ...                   the reference's own CPU path (oracle/_ref)
One "step" = one pass of the whole path over one synthetic recording: by default the 10 MHz cf32 60 s one
of BASELINE.json configs[1] (every rank of an N-GPU run gets its own = configs[4]); --config 3 is the 12 MHz
ci16 / 16384-pt recording of configs[2], --config 4 the 672-burst dense file of configs[3].
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
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")    # one hardware queue per stream (the library itself keeps to 8)

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


def make_recording_gpu(torch, synth, seed, duration_s, bursts_per_s, device, fs=FS, nfft=8192):
    """Config-2/3 style recording generated on the GPU (noise) + CPU-made burst waveforms.  Returns float32 [n, 2]."""
    n = int(duration_s * fs)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    iq = torch.empty((n, 2), dtype=torch.float32, device=device)
    step = 1 << 26
    for o in range(0, n, step):
        m = min(step, n - o)
        iq[o:o + m].normal_(0.0, 0.01, generator=g)
    rng = np.random.default_rng(seed)
    up = fs // 250_000
    pool = [synth.burst_waveform(rng, up, 179, False) for _ in range(48)]
    pool_dev = [(torch.from_numpy(np.stack([w.real, w.imag], 1).astype(np.float32)).to(device), b) for w, b in pool]
    raster = 1e6 / 24.0
    kmax = int((fs / 2 - 200e3) / raster)
    chans = [k * raster for k in range(-kmax, kmax + 1) if abs(k * raster) >= 60e3]
    lead = 512 * nfft / fs + 0.02
    blen = len(pool[0][0])
    nb = int(bursts_per_s * (duration_s - lead))
    starts = np.sort(rng.uniform(lead, duration_s - blen / fs - 0.03, nb))
    last = {}
    truth = []
    kk = torch.arange(blen, device=device, dtype=torch.float64)
    for t0 in starts:
        for _ in range(32):
            ch = chans[int(rng.integers(0, len(chans)))]
            if t0 - last.get(ch, -1.0) >= 0.027 + blen / fs:
                break
        else:
            continue
        last[ch] = float(t0)
        w, bits = pool_dev[int(rng.integers(0, len(pool_dev)))]
        snr = float(rng.uniform(12.0, 25.0))
        amp = 0.01 * 10.0 ** (snr / 20.0)
        f = ch + float(rng.uniform(-3e3, 3e3))
        ph = float(rng.uniform(0, 2 * np.pi))
        s0 = int(round(t0 * fs))
        arg = (2 * np.pi * f / fs) * kk + ph
        c, s = torch.cos(arg).float(), torch.sin(arg).float()
        seg = iq[s0:s0 + blen]
        seg[:, 0] += amp * (w[:, 0] * c - w[:, 1] * s)
        seg[:, 1] += amp * (w[:, 0] * s + w[:, 1] * c)
        truth.append(bits)
    return iq, truth


# BASELINE.json configs[1..4] (configs[0] is the reference's own CPU plumbing case)
WORKLOADS = {
    2: dict(fs=10_000_000, nfft=8192, fmt="cf32", seconds=60.0,
            name="synthetic 10 MHz cf32 recording, 60 s, 8192-pt detect FFT, ~100 bursts/s, full path (BASELINE configs[1])"),
    3: dict(fs=12_000_000, nfft=16384, fmt="ci16", seconds=60.0,
            name="12 MHz ci16 extended-band recording, 60 s, 16384-pt detect FFT, Gardner on, ~100 bursts/s (BASELINE configs[2])"),
    4: dict(fs=10_000_000, nfft=8192, fmt="cf32", seconds=0.85,
            name="high-burst-density synthetic: 672 bursts inside 150 ms of a 0.85 s 10 MHz cf32 file (BASELINE configs[3])"),
    5: dict(fs=10_000_000, nfft=8192, fmt="cf32", seconds=60.0,
            name="independent 10 MHz cf32 streams, 60 s each, one per GPU, no NCCL on the data path (BASELINE configs[4])"),
}


def workload_name(cfg, seconds, bursts_per_s):
    w = WORKLOADS[cfg]
    nm = w["name"]
    if cfg != 4 and (seconds != w["seconds"] or bursts_per_s != 100.0):
        nm += f" [run with {seconds:g} s, ~{bursts_per_s:g} bursts/s]"
    return nm


def build_workload(torch, synth, cfg, seed, seconds, bursts_per_s, dev):
    """-> (device tensor in the file's own format, fmt, fs, n_samples, set of planted bit strings)"""
    w = WORKLOADS[cfg]
    if cfg == 4:
        rec = synth.make_dense_recording(1234)
        t = torch.from_numpy(rec.iq.view(np.float32).reshape(-1, 2)).to(dev)
        return t, "cf32", w["fs"], rec.n_samples, {x.bits for x in rec.truth}
    iq, truth = make_recording_gpu(torch, synth, seed, seconds, bursts_per_s, dev, fs=w["fs"], nfft=w["nfft"])
    if w["fmt"] == "ci16":
        # the reference keeps only the upper byte (main.c:245-246): scale so that it carries the signal
        q = torch.empty(iq.shape, dtype=torch.int16, device=dev)
        step = 1 << 26
        for o in range(0, iq.shape[0], step):
            q[o:o + step] = torch.round(torch.clamp(iq[o:o + step] * (32768.0 * 4.0), -32767.0, 32767.0)).to(torch.int16)
        del iq
        return q, "ci16", w["fs"], q.shape[0], set(truth)
    return iq, "cf32", w["fs"], iq.shape[0], set(truth)


def bind_rank(local, world):
    """CPU affinity (and, where the platform says which, memory policy) of this rank before any pinned allocation:
    the NUMA node of its GPU when sysfs names one, else an equal share of the visible cores."""
    info = {"numa_node": None, "cpus": None}
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local), "pci_domain_id", 0)
        devid = getattr(torch.cuda.get_device_properties(local), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{devid:02x}.0/numa_node"
        node = int(open(path).read().strip()) if os.path.exists(path) else -1
    except Exception:
        node = -1
    try:
        allc = sorted(os.sched_getaffinity(0))
        cpus = None
        if node >= 0 and os.path.exists(f"/sys/devices/system/node/node{node}/cpulist"):
            cpus = []
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus += list(range(int(lo), int(hi or lo) + 1))
            cpus = [c for c in cpus if c in allc] or None
            try:
                import ctypes
                libc = ctypes.CDLL(None, use_errno=True)
                mask = ctypes.c_ulong(1 << node)
                libc.syscall(238, 2, ctypes.byref(mask), ctypes.c_ulong(64))      # set_mempolicy(MPOL_BIND, {node})
            except Exception:
                pass
        if cpus is None and world > 1:
            per = max(len(allc) // world, 1)
            cpus = allc[local * per:(local + 1) * per] or allc
        if cpus:
            os.sched_setaffinity(0, cpus)
        info = {"numa_node": node, "cpus": f"{min(cpus)}-{max(cpus)}" if cpus else "all"}
    except Exception:
        pass
    return info


def kernel_traffic():
    """DRAM bytes per unit of work of every kernel, from the committed ncu captures (profiles/r2_kernel_traffic.json,
    written by tools/ncu_traffic.py out of `ncu --set full` reports of this build)."""
    p = os.path.join(ROOT, "profiles", "r2_kernel_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


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
    cfg = args.config if args.config else (5 if world > 1 else 2)
    W = WORKLOADS[cfg]
    binding = bind_rank(local, world)          # before any pinned allocation (host buffer, the pipeline's arenas)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    synth = importlib.import_module("iridium-sniffer_b200.synth")
    pl = importlib.import_module("iridium-sniffer_b200.pipeline")

    t_gen = time.time()
    seconds = W["seconds"] if cfg == 4 else args.seconds
    iq_dev, fmt, fs, n, truth_set = build_workload(torch, synth, cfg, shard_seed(world, rank), seconds, args.bursts_per_s, dev)
    bps = iq_dev.element_size() * 2
    host = torch.empty(iq_dev.shape, dtype=iq_dev.dtype, pin_memory=True)
    host.copy_(iq_dev)
    torch.cuda.synchronize()
    t_gen = time.time() - t_gen

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    iq_ptr, host_ptr = iq_dev.data_ptr(), host.data_ptr()
    # ---------------- cold run: a fresh pipeline handed the file once (allocations, NCO tables of every channel)
    p = pl.Pipeline(sample_rate=fs, device=local, start_time_ns=1_700_000_000_000_000_000)
    barrier()
    tc = time.perf_counter()
    p.run_host_raw(host_ptr, n, fmt)
    p.raw_text_len("b200")
    cold_ms = (time.perf_counter() - tc) * 1e3

    # ---------------- device-resident leg (value)
    # The timed loops call the C ABI only (run + per-step stats struct); result conversion to
    # Python objects happens once, after the clocks stop.
    for _ in range(args.warmup):
        p.run_device_raw(iq_ptr, n, fmt)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms, launches, stage = 0.0, 0, {}
    for _ in range(args.steps):
        p.run_device_raw(iq_ptr, n, fmt)
        st = p.stats()
        dev_ms += st["ms_total"]
        launches += st["kernel_launches"]
        for k in ("ms_detect_fft", "ms_detect_scan", "ms_downmix_fir", "ms_downmix_chain", "ms_demod"):
            stage[k] = stage.get(k, 0.0) + st[k]
    barrier()
    wall = time.perf_counter() - t0
    res = p.results()
    scan = p.scan_stats()
    # ---------------- end-to-end leg: pinned host IQ -> RAW text lines
    for _ in range(args.warmup):
        p.run_host_raw(host_ptr, n, fmt)
        p.raw_text_len("b200")
    barrier()
    t1 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(args.steps):
        p.run_host_raw(host_ptr, n, fmt)
        p.raw_text_len("b200")                  # every RAW: line of the step, in the library's text buffer
        st = p.stats()
        h2d += st["h2d_bytes"]; d2h += st["d2h_bytes"]
    barrier()
    wall_e2e = time.perf_counter() - t1
    text = bytes(p.raw_text_view())
    n_lines = text.count(b"\n")
    assert text.startswith(b"RAW: b200 ") and n_lines == len(p.results().frames)
    clocks = sampler.stop() if rank == 0 else None
    # context for e2e: the bare pinned->device copy of one step's input, (a) every rank at once between two
    # barriers -- the floor of THIS box at THIS N -- and (b) rank by rank with the others idle
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    conc = []
    for _ in range(3):
        barrier()
        ev0.record()
        iq_dev.copy_(host, non_blocking=True)
        ev1.record()
        torch.cuda.synchronize()
        conc.append(ev0.elapsed_time(ev1))
    barrier()
    alone = []
    for r in range(world):
        if r == rank:
            for _ in range(2):
                torch.cuda.synchronize()
                ev0.record()
                iq_dev.copy_(host, non_blocking=True)
                ev1.record()
                torch.cuda.synchronize()
                alone.append(ev0.elapsed_time(ev1))
        barrier()
    h2d_conc_ms, h2d_alone_ms = min(conc), min(alone)

    (wall, wall_e2e, dev_s, h2d_conc_ms, h2d_alone_max, cold_ms), (n_bursts, n_frames, launches) = reduce_over_ranks(
        torch, dist if world > 1 else None, dev, [wall, wall_e2e, dev_ms / 1e3, h2d_conc_ms, h2d_alone_ms, cold_ms],
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
        tile = 240 if fs // 250_000 == 48 else 256          # IR_FIR_TILE_OF(dec)
        n_tiles = sum((b["dec_len"] + tile - 1) // tile for b in res.bursts if b["dec_len"] >= 100)
        n_det_frames = n // W["nfft"]
        kern = {   # name: (algorithmic bytes, ms of stream time per step, units one step launches, unit name)
            "k_detect_fft": ((bps + 4.0) * n, stage["ms_detect_fft"] / K, n_det_frames, "detector frames"),
            "k_detect_scan": (4.0 * n, stage["ms_detect_scan"] / K, n_det_frames, "detector frames"),   # bitmap pass + state machine
            "k_fir": (bps * sum_n + 8.0 * sum_dec, stage["ms_downmix_fir"] / K, n_tiles, f"{tile}-output tiles"),
            "k_chain": (8.0 * (2 * sum_dec + sum_fl), stage["ms_downmix_chain"] / K, len(res.bursts), "bursts"),
            "k_demod": (8.0 * sum_fl + 5.0 * sum(f["n_bits"] for f in res.frames), stage["ms_demod"] / K, len(res.bursts), "bursts"),
        }
        traffic = kernel_traffic()
        # the kernel the roofline is reported for: the one with the largest share of the machine.  k_fir / k_chain /
        # k_detect_fft launch enough CTAs to fill every SM; the state machine (a few hundred single warps) and the
        # slicer (one thread per frame: ~45 warps per wave) are latency chains whose stream time overlaps everything
        # else -- their ms figures are listed, but "dominant" is decided among the kernels that own the SMs
        wide = [k for k in ("k_detect_fft", "k_fir", "k_chain") if kern[k][1] > 0] or list(kern)
        dom = max(wide, key=lambda k: kern[k][1])
        if kern["k_detect_scan"][1] > 2.0 * kern[dom][1]:
            dom = "k_detect_scan"                    # (a chunk handed to the cluster kernel: the state machine IS the step)
        ach = kern[dom][0] / (kern[dom][1] * 1e-3) / 1e9 if kern[dom][1] > 0 else 0.0
        tr = traffic.get(dom)
        roof = {"bound": "hbm", "kernel": dom, "achieved": round(ach, 2), "peak": peaks["hbm_gbs"],
                "peak_kind": peak_kind + (" burst copy bandwidth" if peak_kind == "measured" else ""),
                "unit": "GB/s", "frac": round(ach / peaks["hbm_gbs"], 5),
                "traffic": round(tr["dram_bytes_per_unit"] * kern[dom][2]) if tr else None,
                "traffic_note": (f"{tr['dram_bytes_per_unit']:.0f} DRAM bytes per {tr['unit']} (ncu --set full, {tr['source']}) x "
                                 f"{kern[dom][2]} {kern[dom][3]} of one step" if tr else "no ncu capture of this kernel committed"),
                "ms_per_step": round(kern[dom][1], 4),
                "how_chosen": "largest stream time among the kernels whose grids fill the machine (k_detect_fft, k_fir, k_chain); "
                              "k_detect_scan only when it exceeds twice that (a chunk handed to the cluster kernel); k_demod is "
                              "one thread per frame (~45 warps a wave), latency overlapped with later waves",
                "whole_path": {"alg_bytes": res.stats["alg_bytes"],
                               "achieved": round(res.stats["alg_bytes"] / (dev_s / K) / 1e9, 2),
                               "frac": round(res.stats["alg_bytes"] / (dev_s / K) / 1e9 / peaks["hbm_gbs"], 5)},
                "kernels": {k: {"alg_gb": round(v[0] / 1e9, 4), "ms": round(v[1], 4),
                                "gbs": round(v[0] / (v[1] * 1e-3) / 1e9, 1) if v[1] > 0 else None}
                            for k, v in kern.items()}}
        if dom == "k_fir":
            # the decimating FIR sits at the fp32 ridge: report the FMA roof beside the HBM one
            flops = 2.0 * 2.0 * 801 * sum_dec
            roof["fp32"] = {"achieved_tflops": round(flops / (kern[dom][1] * 1e-3) / 1e12, 2), "peak_tflops": 74.4,
                            "frac": round(flops / (kern[dom][1] * 1e-3) / 1e12 / 74.4, 4),
                            "peak_kind": "148 SMs x 128 FMA/clk x 1.965 GHz"}
        host_np = host.numpy()
        cpu = cpu_baseline_port(host_np, fmt, fs, args.cpu_seconds)
        out = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": args.warmup, "ms_per_step": round(wall / K * 1e3, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(cfg, seconds, args.bursts_per_s), "baseline_config": cfg,
                       "format": fmt, "sample_rate": fs, "samples_per_gpu": n, "bursts": n_bursts, "raw_frames": n_frames,
                       "bits_matching_ground_truth": f"{ok_bits}/{len(res.frames)} (rank 0)",
                       "l2": "inputs larger than L2 (no flush needed)" if n * bps > (256 << 20) else
                             "input smaller than L2: every step re-reads it from HBM-resident memory after the previous step's "
                             "intermediates (magnitudes, bitmaps: > L2) have passed through",
                       "device_ms_per_step": round(dev_s / K * 1e3, 3), "gen_s": round(t_gen, 1),
                       "state_machine": {"mode": "segmented" if scan.get("segmented") else ("streaming" if scan.get("streaming") else "cluster"),
                                         "chunks_kept": scan["launches_kept"], "chunks_handed_over": scan["launches_bailed"],
                                         "rounds": scan["commands"], "bitmap_rebuilds": scan.get("bitmap_rebuilds", 0),
                                         "last_hand_over_reason": scan.get("last_bail_reason", 0),
                                         "crowded_segment_walks": scan.get("generic_segment_walks", 0)}},
            "e2e": {"value": round(total / wall_e2e / 1e6, 2), "unit": UNIT,
                    "h2d_bytes_per_step": h2d // K, "d2h_bytes_per_step": d2h // K,
                    "ms_per_step": round(wall_e2e / K * 1e3, 3), "raw_lines_per_step": n_lines,
                    "h2d_concurrent_ms": round(h2d_conc_ms, 3),
                    "h2d_concurrent_gbs_per_gpu": round(n * bps / h2d_conc_ms / 1e6, 1),
                    "h2d_copy_alone_ms": round(h2d_alone_max, 3),
                    "h2d_copy_alone_gbs": round(n * bps / h2d_alone_max / 1e6, 1),
                    "frac_of_concurrent_copy_floor": round(h2d_conc_ms / (wall_e2e / K * 1e3), 3),
                    "cold_first_step_ms": round(cold_ms, 1),
                    "binding": binding,
                    "api": "ir_pipeline_run_host (pinned host IQ -> frames) + ir_pipeline_format_raw_all"},
            "bursts_per_s": round(n_bursts / (wall / K), 1),
            "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(out))
    p.close()
    if world > 1:
        dist.destroy_process_group()


def _host_cf32(host_np, fmt, m):
    """first m samples of the host buffer as complex64 (ci16 the way the reference's file reader does it)"""
    if fmt == "cf32":
        return host_np[:m].view(np.complex64).reshape(-1)
    from oracle import bindings as ob
    return ob.Port().convert_ci16(np.ascontiguousarray(host_np[:m]).reshape(-1))


def cpu_baseline_port(host_np, fmt, fs, seconds):
    """Oracle restatement (single thread) on the first `seconds` of the same recording."""
    from oracle import bindings as ob
    port = ob.Port()
    m = min(host_np.shape[0], int(seconds * fs))
    iq = _host_cf32(host_np, fmt, m)
    t = time.perf_counter()
    res, st = port.run(iq, sample_rate=fs)
    dt = time.perf_counter() - t
    return {"value": round(m / dt / 1e6, 2), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"first {m / fs:g} s ({m} samples) of the rank-0 recording: "
                      f"{st['n_bursts']} bursts, {st['n_results']} frames, "
                      f"detect {st['t_detect_s']:.2f}s downmix {st['t_downmix_s']:.2f}s demod {st['t_demod_s']:.2f}s"}


def run_reference(args):
    """The reference's own program (oracle/_ref/iridium-sniffer: unmodified sources + FFT shim,
    AVX2 kernels, its fixed 1+4+1 thread graph) on a bounded sample -- a prefix -- of the same recording
    the CUDA arm decodes (same generator, same seed)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import bindings as ob
    synth = importlib.import_module("iridium-sniffer_b200.synth")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = args.config if args.config else (5 if world > 1 else 2)
    W = WORKLOADS[cfg]
    fs = W["fs"]
    if not os.path.exists(ob.REF_BIN):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/iridium-sniffer not built"}))
        return
    secs = min(args.ref_seconds, args.seconds if cfg != 4 else W["seconds"])
    how = "prefix of the CUDA arm's rank-0 recording (same generator, same seed)"
    try:
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("no GPU for the generator")
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        seconds = W["seconds"] if cfg == 4 else args.seconds
        t, fmt, fs, n_all, _ = build_workload(torch, synth, cfg, shard_seed(world, 0), seconds, args.bursts_per_s, dev)
        n = min(n_all, int(secs * fs))
        arr = t[:n].cpu().numpy()
        del t
        torch.cuda.empty_cache()
    except Exception as e:                     # no GPU here: the CPU generator with the same recipe
        how = f"CPU generator, same recipe, seed 2 ({type(e).__name__})"
        fmt = W["fmt"]
        if cfg == 4:
            rec = synth.make_dense_recording(1234)
        else:
            rec = synth.make_recording(2, sample_rate=fs, duration_s=secs, fmt=fmt,
                                       n_bursts=int(args.bursts_per_s * max(secs - 512 * W["nfft"] / fs - 0.05, 0.1)), waveform_pool=32)
        arr = rec.iq if fmt != "cf32" else rec.iq.view(np.float32)
        n = rec.n_samples
    path = ("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp") + f"/ir_bench_ref.{fmt}"
    arr.tofile(path)
    del arr

    def one():
        t = time.perf_counter()
        # stdout/stderr go to files: a pipe nobody drains would block the child after 64 KiB
        fo, fe = open(path + ".out", "wb"), open(path + ".err", "wb")
        pr = subprocess.Popen([ob.REF_BIN, "-f", path, f"--format={fmt}", "-r", str(fs), "--file-info=ref"],
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
    import ctypes.util
    fftw = ctypes.util.find_library("fftw3f")
    out = {
        "impl": "reference", "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(bottleneck * 1e3, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(cfg, W["seconds"] if cfg == 4 else args.seconds, args.bursts_per_s),
                   "baseline_config": cfg, "format": fmt, "sample_rate": fs,
                   "sample": f"first {n / fs:g} s ({n} samples): {how}",
                   "timing": "samples / CPU time of the busiest thread (the detector): the reference's pipeline "
                             "throughput without its 1 s exit quantum; wall-clock figure in cpu_baseline.sample",
                   "raw_lines": nl},
        "cpu_baseline": {"value": round(value, 2), "unit": UNIT, "cores": min(7, cores), "kind": "reference",
                         "sample": f"{n / fs:g} s of signal; threads seen {nth} (1 detector + 4 downmix + 1 demod + reader + stats: "
                                   f"the reference's fixed thread graph); host cores {cores}; "
                                   f"wall {np.mean(walls):.2f} s -> {n / np.mean(walls) / 1e6:.1f} Msps wall; FFT: oracle/shim "
                                   f"radix-2 (the binary was built where libfftw3f is absent; on this box libfftw3f is "
                                   f"{'present: ' + fftw if fftw else 'absent too'}, and /root/reference is not here to rebuild against it)"},
        "e2e": {"value": round(value, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=0, choices=[0, 2, 3, 4, 5],
                    help="BASELINE.json config (1-based): 2 = 10 MHz cf32 60 s (default at N=1), 3 = 12 MHz ci16 60 s, "
                         "4 = 672-burst dense file, 5 = one independent config-2 stream per GPU (default at N>1)")
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
