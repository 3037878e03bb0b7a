"""SURVEY.md 8e (2) on the GPU: one stream cut into time blocks (ir_plan_blocks), each block through the CUDA path
with its origin set (ir_pipeline_set_origin + ir_pipeline_run_host), the frame lists merged (ir_merge_blocks).
Two checks: (1) block by block the CUDA path gives what the CPU oracle gives on the same samples -- the path's usual
bar: same frames, same bits, time stamps on the stream's clock; (2) the merge of the GPU blocks equals the GPU run
over the whole stream by the criteria tests/test_time_blocks.py states (bits exact for every matched burst, >= 99 %
matched, each boundary burst exactly once).

The plan and the merge are host code and are pinned on the CPU against the oracle (tests/test_time_blocks.py); this
file adds the device side -- the origin entering the device path's time stamps, blocks of different sizes through one
pipeline, one process owning several GPUs (the last case needs two devices and skips only when the box has one).
Not collected by name; tests/test_zz_gpu_classify.py runs it case by case in child processes, last."""
import importlib
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(HERE, name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


tb = _load("test_time_blocks")
T0 = tb.T0


def test_time_blocks_through_the_cuda_path():
    pl = importlib.import_module("iridium-sniffer_b200.pipeline")
    synth = importlib.import_module("iridium-sniffer_b200.synth")
    from oracle import bindings as ob
    ob.build(port=True, ref=False)
    port = ob.Port()
    rec, cfg, blocks = tb._recording(synth)
    assert len(blocks) == 3
    p = pl.Pipeline(sample_rate=rec.sample_rate, start_time_ns=T0)
    whole = p.run_host(rec.iq).frames
    per_block = [p.run_block(rec.iq, b).frames for b in blocks]          # one pipeline, blocks of different sizes
    again = p.run_host(rec.iq).frames                                    # the origin does not stick to later runs
    assert [(f["id"], f["timestamp"]) for f in again] == [(f["id"], f["timestamp"]) for f in whole]

    # (1) each block against the oracle on the same samples
    want = tb._oracle_blocks(port, rec, blocks)
    for k, (got, exp) in enumerate(zip(per_block, want)):
        assert len(got) == len(exp) and len(got) > 20, (k, len(got), len(exp))
        for g, e in zip(sorted(got, key=lambda d: d["id"]), sorted(exp, key=lambda d: d["id"])):
            assert g["id"] == e["id"] and tb._bitstr(g) == tb._bitstr(e), (k, g["id"])
            assert abs(int(g["timestamp"]) - int(e["timestamp"])) <= 1, (k, g["id"], g["timestamp"], e["timestamp"])
            assert abs(g["center_frequency"] - e["center_frequency"]) <= 2.0
            assert abs(g["magnitude"] - e["magnitude"]) <= 0.05 and abs(g["noise"] - e["noise"]) <= 0.05

    # (2) the merge against the run over the whole stream
    merged = pl.merge_blocks(cfg, T0, blocks, per_block)
    matched, exact, missing, extra, dmag = tb._compare(whole, merged, blocks, rec.sample_rate)
    assert len(whole) >= 120
    assert matched >= 0.99 * len(whole), (len(whole), matched)
    assert exact == matched
    assert len(extra) <= max(1, len(whole) // 100)
    dm = np.array(dmag)
    assert (dm[:, 2] <= 2).all() and (dm[:, 3] <= 1).all()
    ts = [d["timestamp"] for d in merged]
    assert ts == sorted(ts) and len({d["id"] for d in merged}) == len(merged)
    for b in blocks[:-1]:
        e_ns = T0 + int(b.own_end) * 100
        near_w = [w for w in whole if abs(w["timestamp"] - e_ns) < 15_000_000]
        near_m = [d for d in merged if abs(d["timestamp"] - e_ns) < 15_000_000]
        assert len(near_w) >= 4 and len(near_m) == len(near_w)
    # and the merge of the GPU blocks is the merge of the oracle's blocks, line for line
    want_m = pl.merge_blocks(cfg, T0, blocks, want)
    assert [(d["id"], d["block"], tb._bitstr(d)) for d in merged] == [(d["id"], d["block"], tb._bitstr(d)) for d in want_m]
    p.close()


def test_one_process_driver_on_the_gpu():
    """ir_multi_*: the same three blocks through one C call (one device here, so its thread works through them in
    order), merged and formatted inside the library == the block-by-block route above"""
    pl = importlib.import_module("iridium-sniffer_b200.pipeline")
    synth = importlib.import_module("iridium-sniffer_b200.synth")
    rec, cfg, blocks = tb._recording(synth)
    p = pl.Pipeline(sample_rate=rec.sample_rate, start_time_ns=T0)
    want = pl.merge_blocks(cfg, T0, blocks, [p.run_block(rec.iq, b).frames for b in blocks])
    p.close()
    m = pl.Multi([0], sample_rate=rec.sample_rate, start_time_ns=T0)
    got = m.run_host(rec.iq, "cf32", n_blocks=3)
    assert [(d["id"], d["timestamp"], d["block"], tb._bitstr(d)) for d in got] == \
        [(d["id"], d["timestamp"], d["block"], tb._bitstr(d)) for d in want]
    r = m.results()
    assert r.n_blocks == 3 and r.kernel_launches > 0 and r.samples_fed > rec.n_samples
    lines = m.raw_text("T").decode().splitlines()
    assert len(lines) == len(got) >= 120 and all(l.startswith("RAW: T ") for l in lines)
    assert all(tb._bitstr(d) == l.split()[-1] for d, l in zip(got, lines))
    with pytest.raises(RuntimeError):
        pl.Multi([0, 0], sample_rate=rec.sample_rate)
    with pytest.raises(RuntimeError):
        pl.Multi([4096], sample_rate=rec.sample_rate)
    m.close()


def test_one_process_driver_parsed_on_the_gpu():
    """`--parsed` through ir_multi_*: one block = the single pipeline's lines, byte for byte, in TIME order (the pipeline
    lists frames in emission order: the weak second detection of a strong burst can precede the first); two blocks =
    the same decoded content"""
    pl = importlib.import_module("iridium-sniffer_b200.pipeline")
    synth = importlib.import_module("iridium-sniffer_b200.synth")
    fg = _load("frame_gen")
    rng = np.random.default_rng(3)
    duplex = [fg.make_ida(rng, 9), fg.make_ida(rng, 20), fg.make_ida(rng, 0), fg.make_ida(rng, 13, good_crc=False)]
    rec = synth.make_recording(21, duration_s=1.5, n_bursts=14, snr_db=(22.0, 28.0), frame_bits=duplex)
    p = pl.Pipeline(sample_rate=rec.sample_rate, start_time_ns=T0)
    p.run_host(rec.iq)
    want = p.parsed_text("T").decode()
    p.close()
    assert want.count("IDA: p-") >= 8
    m = pl.Multi([0], sample_rate=rec.sample_rate, start_time_ns=T0)
    m.set_classify(True)
    m.run_host(rec.iq, "cf32", n_blocks=1)
    got1 = m.parsed_text("T").decode().splitlines()
    assert sorted(got1) == sorted(want.splitlines())
    assert [l.split()[2] for l in got1] == sorted((l.split()[2] for l in got1), key=float)
    fr = m.run_host(rec.iq, "cf32", n_blocks=2)
    assert {d["block"] for d in fr} == {0, 1}
    two = m.parsed_text("T").decode().splitlines()
    one = got1
    assert len(two) == len(one)
    # block 0 starts where the stream starts: its lines are the single pipeline's, byte for byte; block 1 has its own
    # noise baseline (level|noise|snr may move in the last digit), the decoded content from LCW( on is the same
    n0 = sum(d["block"] == 0 for d in fr)
    assert n0 >= 2 and all(l in set(one) for d, l in zip(fr, two) if d["block"] == 0)
    assert [l[l.index("LCW("):] for l in one if l.startswith("IDA:")] == [l[l.index("LCW("):] for l in two if l.startswith("IDA:")]
    m.close()


def test_one_process_driver_independent_streams_on_the_gpu():
    """BASELINE config 5 through ir_multi_run_streams_host: each stream's frames == that stream through a pipeline of
    its own (one device here: the streams run one after the other on it)"""
    pl = importlib.import_module("iridium-sniffer_b200.pipeline")
    synth = importlib.import_module("iridium-sniffer_b200.synth")
    recs = [synth.make_recording(40 + s, duration_s=0.62 + 0.05 * s, n_bursts=4 + s, snr_db=(15.0, 22.0)) for s in range(3)]
    want = []
    for r in recs:
        p = pl.Pipeline(sample_rate=r.sample_rate, start_time_ns=T0)
        want.append(p.run_host(r.iq).frames)
        p.close()
    m = pl.Multi([0], sample_rate=recs[0].sample_rate, start_time_ns=T0)
    got = m.run_streams_host([r.iq for r in recs])
    flat = [(s, d) for s, fl in enumerate(want) for d in fl]
    assert len(got) == len(flat) >= 9
    for g, (s, w) in zip(got, flat):
        assert g["block"] == s and g["id"] == s * pl.BLOCK_ID_STRIDE + w["id"] and g["timestamp"] == w["timestamp"]
        assert tb._bitstr(g) == tb._bitstr(w)
    assert len(m.raw_text("T").decode().splitlines()) == len(got)
    m.close()


def test_two_devices_blocks_and_streams():
    """ir_multi_* owning TWO GPUs (north_star: the stream shards by contiguous time blocks across the GPUs of one box,
    independent per-GPU pipelines, results gathered on the host): the blocks dealt to devices 0 and 1 give the merge
    a single device gives, line for line; two independent streams on two devices give what a pipeline of their own
    gives.  Skips only on a one-GPU box."""
    pl = importlib.import_module("iridium-sniffer_b200.pipeline")
    synth = importlib.import_module("iridium-sniffer_b200.synth")
    if pl.load_library().ir_device_count() < 2:
        pytest.skip("needs two CUDA devices")
    rec, cfg, blocks = tb._recording(synth)
    one = pl.Multi([0], sample_rate=rec.sample_rate, start_time_ns=T0)
    want = one.run_host(rec.iq, "cf32", n_blocks=4)
    want_txt = one.raw_text("T")
    one.close()
    two = pl.Multi([0, 1], sample_rate=rec.sample_rate, start_time_ns=T0)
    got = two.run_host(rec.iq, "cf32", n_blocks=4)
    assert len(got) >= 120
    assert [(d["id"], d["timestamp"], d["block"], tb._bitstr(d)) for d in got] == \
        [(d["id"], d["timestamp"], d["block"], tb._bitstr(d)) for d in want]
    assert two.raw_text("T") == want_txt
    r = two.results()
    assert r.n_blocks == 4 and r.samples_fed > rec.n_samples
    # config 5 in one process on two devices
    recs = [synth.make_recording(50 + s, duration_s=0.62 + 0.05 * s, n_bursts=4 + s, snr_db=(15.0, 22.0)) for s in range(4)]
    ref = []
    for q in recs:
        p = pl.Pipeline(sample_rate=q.sample_rate, start_time_ns=T0)
        ref.append(p.run_host(q.iq).frames)
        p.close()
    got = two.run_streams_host([q.iq for q in recs])
    flat = [(s, d) for s, fl in enumerate(ref) for d in fl]
    assert len(got) == len(flat) >= 12
    for g, (s, w) in zip(got, flat):
        assert g["block"] == s and g["id"] == s * pl.BLOCK_ID_STRIDE + w["id"] and g["timestamp"] == w["timestamp"]
        assert tb._bitstr(g) == tb._bitstr(w)
    two.close()
