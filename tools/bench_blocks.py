"""Strong-scaling companion of bench.py: ONE 10 MHz cf32 recording cut into as many time blocks as there are GPUs
(ir_plan_blocks: halo of 512 frames + the longest burst before each block, tail after it), block r through rank r's
pipeline (ir_pipeline_set_origin + ir_pipeline_run_device), frame lists gathered on rank 0 and merged by time stamp
(ir_merge_blocks).  No collective on the data path; torch.distributed carries the barrier, the MAX-over-ranks of the
times and the final gather of the frames.  Rank 0 also runs the whole recording through its own pipeline once and says
how the merge compares (frames matched, bits equal).  Every rank builds the same seeded recording on its own GPU.

Prints one JSON line in bench.py's vocabulary with "scaling": "strong".  A developer aid for profiles/ (measured lines:
profiles/r2*_blocks_strong_n*.json, table in DESIGN.md section 7), not the driver's bench contract.

    python tools/bench_blocks.py [--seconds 60] [--steps 5] [--warmup 3]                         # one GPU, one block
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \\
        tools/bench_blocks.py --seconds 60"""
import argparse
import importlib
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

T0 = 1_700_000_000_000_000_000


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--bursts-per-s", type=float, default=100.0)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    synth = importlib.import_module("iridium-sniffer_b200.synth")
    pl = importlib.import_module("iridium-sniffer_b200.pipeline")

    iq, truth = bench.make_recording_gpu(torch, synth, 2, args.seconds, args.bursts_per_s, dev)   # same on every rank
    n = iq.shape[0]
    cfg = pl.make_config(sample_rate=bench.FS)
    blocks = pl.plan_blocks(cfg, n, world)
    p = pl.Pipeline(sample_rate=bench.FS, device=local, start_time_ns=T0)
    mine = blocks[rank] if rank < len(blocks) else None       # a stream too short for `world` blocks leaves ranks idle
    if mine is not None:
        ff, fe = int(mine.feed_first), int(mine.feed_end)
        ptr, m = iq.data_ptr() + ff * 8, fe - ff
        p.set_origin(ff)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        if mine is not None:
            p.run_device_raw(ptr, m, "cf32")

    for _ in range(args.warmup):
        step()
    barrier()
    t = time.perf_counter()
    launches = 0
    for _ in range(args.steps):
        step()
        launches += p.stats()["kernel_launches"] if mine is not None else 0
    barrier()
    wall = time.perf_counter() - t
    frames = p.results().frames if mine is not None else []
    (wall,), (launches, fed) = bench.reduce_over_ranks(torch, dist if world > 1 else None, dev, [wall],
                                                       [launches, (fe - ff) if mine is not None else 0])
    payload = [{k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in f.items() if k != "llr"} for f in frames]
    gathered = [None] * world if rank == 0 else None
    if world > 1:
        dist.gather_object(payload, gathered, dst=0)
    else:
        gathered = [payload]
    if rank == 0:
        lists = [gathered[k] if k < world else [] for k in range(len(blocks))]
        merged = pl.merge_blocks(cfg, T0, blocks, lists)
        p.set_origin(0)
        whole = p.run_device_ptr(iq.data_ptr(), n, "cf32").frames
        key = lambda d: (round(d["center_frequency"] / 500.0), d["timestamp"] // 1_000_000)
        by = {}
        for d in merged:
            by.setdefault(key(d), []).append(d)
        matched = exact = 0
        for w in whole:
            kf, kt = key(w)
            cand = [d for a in (-1, 0, 1) for b in (-1, 0, 1) for d in by.get((kf + a, kt + b), [])
                    if abs(d["timestamp"] - w["timestamp"]) < 1_000_000 and abs(d["center_frequency"] - w["center_frequency"]) < 200]
            if cand:
                matched += 1
                exact += any(list(c["bits"]) == list(w["bits"]) for c in cand)
        ts = set(truth)
        ok = sum("".join(map(str, d["bits"])) in ts for d in merged)
        K = args.steps
        print(json.dumps({
            "metric": bench.METRIC, "value": round(n * K / wall / 1e6, 2), "unit": bench.UNIT, "n_gpus": world, "steps": K,
            "warmup": args.warmup, "ms_per_step": round(wall / K * 1e3, 3), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"ONE synthetic 10 MHz cf32 recording of {args.seconds:g} s ({n} samples) in {len(blocks)} time "
                                   f"blocks over {world}xB200 (halo {int(blocks[-1].own_first - blocks[-1].feed_first)} samples, "
                                   "device-resident, host merge, no NCCL on the data path)",
                       "samples_fed_per_step": fed, "overhead": round(fed / n - 1.0, 4),
                       "merged_frames": len(merged), "unsharded_frames": len(whole), "matched": matched, "bits_equal": exact,
                       "bits_matching_ground_truth": f"{ok}/{len(merged)}",
                       "l2": "inputs larger than L2 (no flush needed)"},
            "gpu_launches": launches}))
    p.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
