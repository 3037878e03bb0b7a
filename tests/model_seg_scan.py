"""Executable model (numpy, CPU) of the SEGMENTED detector state machine of
iridium-sniffer_b200/csrc/k_detect_seg.cu -- the ALGORITHM, not the kernels.

The reference's frame loop (burst_detect.c:426-632) is serial through two things: the list of active
bursts and the noise baseline (updated on quiet frames only, :438-454).  The segmented scheme breaks
both by speculation and proves the result by a fixed point:

  * a chunk of frames is cut every SEG frames; every segment is walked on its own, from the burst
    list the PREVIOUS round left at its first frame (round 0: empty);
  * the baseline every exact test needs is B_v, "the baseline after v quiet frames of this chunk",
    computed bin-parallel from the quiet flags of the PREVIOUS round (round 0: no frame quiet) with
    the reference's two roundings per update, and kept ("snapshot") for the versions some frame with
    a set bitmap bit can ask for;
  * a round whose outputs (quiet flags, burst lists at the cuts, creation counts) equal the previous
    round's outputs is exact: by induction over the frames, every decision was taken from the true
    burst list and the true baseline.  Ids are (segment, ordinal) codes until the end, when a prefix
    sum over the segments' creation counts gives the reference's running counter.

Guard-banded bitmaps (X = certainly above, XU = above or uncertain for any baseline within the bin's
band, at first [0.65, 1.5] x the chunk's first baseline) are used exactly as in the streaming kernel:
they decide what they prove, everything else is an IEEE divide on B_v.  A baseline that leaves its
band (a channel that carried a burst while the detector primed: its inflated baseline collapses when
those rows rotate out of the history) gets a band that covers what it did, the bitmaps are rebuilt and
the rounds go on.  Squelch, a burst that may exceed max_burst_len, a baseline that is not positive and
finite and a 33rd concurrent burst make the chunk "bail" (the cluster kernel redoes it); a bail that
only a wrong speculative start produced disappears in the next round.

tests/test_seg_scan_model.py runs this on the CPU oracle's magnitude frames and demands the oracle's
burst list field for field."""
import numpy as np

GUARD_LO = np.float32(0.65)
GUARD_HI = np.float32(1.5)
MAX_LANES = 32
BIG = 0x3fffffff


class Bail(Exception):
    def __init__(self, reason, frame=-1):
        super().__init__(f"bail: {reason} at frame {frame}")
        self.reason, self.frame = reason, frame


class SegScanModel:
    def __init__(self, N, thr, half_bw, pre_len, post_len, max_burst_len, max_bursts, hist_size=512, seg=64,
                 max_rounds=12, c_walker=None, lanes=MAX_LANES, walker_lanes=1):
        """c_walker: ctypes handle of tests/seg_generic_host_shim.cpp -- the product's generic segment walker
        (csrc/seg_generic.cuh) compiled for the host; when given, every segment is walked by IT instead of the numpy
        walker below (segment length must be the library's IR_SEG_LEN), with `lanes` as the burst capacity."""
        self.cw, self.lanes = c_walker, lanes
        self.walker_lanes = walker_lanes      # 1, or 3 / 4: that many threads play the warp's lanes
        self.N, self.thr, self.half_bw = N, np.float32(thr), half_bw
        self.pre, self.post, self.max_len, self.max_bursts, self.H = pre_len, post_len, max_burst_len, max_bursts, hist_size
        self.SEG, self.max_rounds = seg, max_rounds
        self.base = np.zeros(N, np.float32)
        self.hist = np.zeros((hist_size, N), np.float32)
        self.hist_idx, self.primed = 0, False
        self.act = []                      # dicts: id, start, last_active, cb, rel, basec
        self.next_id, self.index, self.sq = 0, 0, 0
        self.gone = []
        bins = np.arange(N)
        self.valid = (bins >= half_bw) & (bins < N - half_bw) & ~((bins >= N // 2 - 3) & (bins <= N // 2 + 3))
        self.stats = dict(chunks=0, rounds=0, events=0, snapshots=0, max_rounds_seen=0)

    # ------------------------------------------------------------------ priming (every frame quiet)
    def _prime(self, mag):
        for f in range(mag.shape[0]):
            old = self.hist[self.hist_idx] if self.primed else np.zeros(self.N, np.float32)
            t = self.base - old
            self.base = t + mag[f]
            self.hist[self.hist_idx] = mag[f]
            self.hist_idx += 1
            if self.hist_idx == self.H:
                self.primed, self.hist_idx = True, 0
                if f + 1 < mag.shape[0]:
                    raise Bail("primed in mid-launch", f)

    # ------------------------------------------------------------------ baseline pass of one round
    def _base_pass(self, mag, Q, rowany, guard):
        """B_v for the versions a frame with bitmap bits may ask for; final baseline; per-bin extremes of the baseline."""
        F = mag.shape[0]
        qlist = np.nonzero(Q)[0]
        ver = np.concatenate(([0], np.cumsum(Q)[:-1])).astype(np.int64) if F else np.zeros(0, np.int64)
        vneed = np.zeros(len(qlist) + 1, bool)
        vneed[ver[rowany]] = True
        snaps = {}
        base = self.base.copy()
        bmin, bmax = base.copy(), base.copy()
        if vneed[0]:
            snaps[0] = base.copy()
        for q, f in enumerate(qlist):
            # the q-th quiet frame of the chunk overwrites history row (hist_idx + q) % H: what it
            # replaces is the carried-in history for q < H, else the chunk's own quiet frame q - H
            old = self.hist[(self.hist_idx + q) % self.H] if q < self.H else mag[qlist[q - self.H]]
            t = base - old
            base = t + mag[f]
            bmin, bmax = np.minimum(bmin, base), np.maximum(bmax, base)
            if vneed[q + 1]:
                snaps[q + 1] = base.copy()
        self.stats["snapshots"] += len(snaps)
        return ver, snaps, base, qlist, (bmin, bmax)

    def _free_mask(self, act):
        free = np.ones(self.N, bool)
        for b in act:
            free[max(b["cb"] - self.half_bw, 0):min(b["cb"] + self.half_bw, self.N - 1) + 1] = False
        return free

    # ------------------------------------------------------------------ one segment
    def _walk(self, s, f0, f1, start, mag, X, XU, ver, snaps, index0, sq0):
        """frames [f0, f1) from the burst list `start`; returns (end list, quiet flags, gone, n_create, bail)"""
        if self.cw is not None:
            return self._walk_c(s, f0, f1, start, mag, X, XU, ver, snaps, index0, sq0)
        N, thr = self.N, self.thr
        PF = -(-self.post // N)
        PF0 = max(1, -(-(self.post - self.pre) // N))
        TLF = (self.max_len - self.pre) // N if self.max_len >= self.pre else -1
        if self.max_len <= 0:
            TLF = BIG
        act = [dict(b) for b in start]
        if len(act) > MAX_LANES:
            return None, None, None, 0, "more bursts than lanes"
        q = np.zeros(f1 - f0, np.uint8)
        gone, ncreate = [], 0
        sq = max(sq0 - f0, 0)
        free = self._free_mask(act)
        for f in range(f0, f1):
            idx = index0 + f * N
            fv = free & self.valid
            ev = bool(np.any(XU[f] & fv))
            for bst in act:
                cb = bst["cb"]
                bst["x3"] = bool(X[f, max(cb - 1, 0):cb + 2].any())
                bst["u3"] = bool(XU[f, max(cb - 1, 0):cb + 2].any())
                if (not bst["x3"] and (bst["u3"] or f >= bst["dl"])) or f > bst["tl"]:
                    ev = True
            if not ev:
                for bst in act:
                    if bst["x3"]:
                        bst["dl"], bst["lah"] = f + PF, f
                sq = max(sq - 1, 0)
                if not act:
                    q[f - f0] = 1
                continue
            self.stats["events"] += 1
            if any(f > bst["tl"] for bst in act):
                return None, None, None, 0, "too long"
            row = mag[f]
            B = snaps.get(int(ver[f]))          # exact baseline of this frame (None: no bit set in the row)
            for bst in act:
                hit = bst["x3"]
                if not hit and bst["u3"]:
                    for bn in range(max(bst["cb"] - 1, 0), min(bst["cb"] + 2, N)):
                        if XU[f, bn] and B[bn] > 0 and np.float32(row[bn] / B[bn]) > thr:
                            hit = True
                if hit:
                    bst["dl"], bst["lah"] = f + PF, f
                bst["done"] = (not hit) and f >= bst["dl"]
            peaks = []
            for bn in np.nonzero(XU[f] & fv)[0]:
                if B[bn] > 0:
                    rel = np.float32(row[bn] / B[bn])
                    if rel > thr:
                        peaks.append((rel, int(bn)))
            for bst in sorted([b for b in act if b["done"]], key=lambda b: b["id"]):
                la = bst["last_active"] if bst["lah"] is None else index0 + bst["lah"] * N
                gone.append((bst["id"], bst["start"], idx, la, bst["cb"], bst["rel"], bst["basec"]))
            if any(b["done"] for b in act):
                act = [b for b in act if not b["done"]]
                free = self._free_mask(act)
            peaks.sort(key=lambda p: (-p[0], p[1]))
            for rel, bn in peaks:
                if not free[bn]:
                    continue
                if len(act) >= MAX_LANES:
                    return None, None, None, 0, "a 33rd concurrent burst"
                st = idx - self.pre
                act.append(dict(id=(1, s, ncreate), start=st, last_active=st, cb=bn, rel=rel, basec=B[bn],
                                dl=f + PF0, lah=None, tl=f + TLF))
                ncreate += 1
                free[max(bn - self.half_bw, 0):min(bn + self.half_bw, N - 1) + 1] = False
            if self.max_bursts > 0 and len(act) > self.max_bursts:
                return None, None, None, 0, "squelch"
            if sq > 0:
                sq -= 1
            if not act:
                q[f - f0] = 1
        for bst in act:
            bst.pop("x3", None); bst.pop("u3", None); bst.pop("done", None)
        return act, q, gone, ncreate, None

    # ------------------------------------------------------------------ one segment, by the product's C++ walker
    def _walk_c(self, s, f0, f1, start, mag, X, XU, ver, snaps, index0, sq0):
        import ctypes as C
        N, W = self.N, self.N // 32
        key = (id(X), id(XU))
        if getattr(self, "_pack_key", None) != key:            # bitmaps of the chunk in the kernels' layout: [XU words][X words]
            xu = np.empty((X.shape[0], 2 * W), np.uint32)
            xu[:, :W] = np.packbits(XU, axis=1, bitorder="little").view(np.uint32)
            xu[:, W:] = np.packbits(X, axis=1, bitorder="little").view(np.uint32)
            self._xu, self._pack_key = np.ascontiguousarray(xu), key
            self._valid_w = np.packbits(self.valid, bitorder="little").view(np.uint32).copy()
        skey = id(snaps)
        if getattr(self, "_snap_key", None) != skey:           # snapshots as [slots][N] + a slot per frame
            vers = sorted(snaps)
            self._snap = np.ascontiguousarray(np.stack([snaps[v] for v in vers]) if vers else np.zeros((1, N), np.float32), np.float32)
            slot_of = {v: i for i, v in enumerate(vers)}
            self._fslot = np.array([slot_of.get(int(v), -1) for v in ver], np.int32)
            self._snap_key = skey
        CODE = 1 << 63
        cap = self.lanes
        Burst = np.dtype([("id", "<u8"), ("start", "<u8"), ("last0", "<u8"), ("cb", "<i4"), ("rel", "<f4"), ("base", "<f4"),
                          ("dl", "<i4"), ("lah", "<i4"), ("tl", "<i4")])
        Gone = np.dtype([("id", "<u8"), ("start", "<u8"), ("stop", "<u8"), ("la", "<u8"), ("cb", "<i4"), ("rel", "<f4"),
                         ("base", "<f4"), ("pad", "<i4")])
        assert Burst.itemsize == self.cw.segg_sizeof_burst() and Gone.itemsize == self.cw.segg_sizeof_gone()
        if len(start) > cap:
            return None, None, None, 0, "more bursts than lanes"
        work = np.zeros(cap + 1, Burst)
        NONE = -0x40000000

        def enc(code):
            return code[1] if code[0] == 0 else CODE | (code[1] << 32) | code[2]

        def dec(v):
            v = int(v)
            return (1, (v >> 32) & 0x7fffffff, v & 0xffffffff) if v & CODE else (0, v, 0)
        for i, b in enumerate(start):
            work[i] = (enc(b["id"]), b["start"], b["last_active"], b["cb"], b["rel"], b["basec"], b["dl"],
                       NONE if b["lah"] is None else b["lah"], b["tl"])
        gl = np.zeros(4096, Gone)
        counts = (C.c_int * 3)()
        qb = (C.c_uint32 * 8)()
        rc = self.cw.segg_walk(N, self.half_bw, self.max_bursts, self.pre, self.post, self.max_len, C.c_float(float(self.thr)),
                               s, f0, f1 - f0, C.c_longlong(index0 + f0 * N), max(sq0 - f0, 0),
                               self._xu.ctypes.data_as(C.c_void_p), np.ascontiguousarray(mag, np.float32).ctypes.data_as(C.c_void_p),
                               self._snap.ctypes.data_as(C.c_void_p), self._fslot.ctypes.data_as(C.c_void_p),
                               self._valid_w.ctypes.data_as(C.c_void_p), work.ctypes.data_as(C.c_void_p), len(start), cap,
                               gl.ctypes.data_as(C.c_void_p), len(gl), counts, qb, self.walker_lanes)
        if rc != 0:
            assert rc > 0, f"the walker's lanes disagree ({rc})"
            return None, None, None, 0, {3: "too long", 4: "peak list", 5: "a 33rd concurrent burst", 6: "squelch",
                                         9: "missing snapshot", 10: "gone list"}[rc]
        n_end, n_gone, n_create = counts[0], counts[1], counts[2]
        act = [dict(id=dec(w["id"]), start=int(w["start"]), last_active=int(w["last0"]), cb=int(w["cb"]), rel=np.float32(w["rel"]),
                    basec=np.float32(w["base"]), dl=int(w["dl"]), lah=None if int(w["lah"]) == NONE else int(w["lah"]), tl=int(w["tl"]))
               for w in work[:n_end]]
        q = np.array([(qb[k >> 5] >> (k & 31)) & 1 for k in range(f1 - f0)], np.uint8)
        gone = [(dec(g["id"]), int(g["start"]), int(g["stop"]), int(g["la"]), int(g["cb"]), np.float32(g["rel"]), np.float32(g["base"]))
                for g in gl[:n_gone]]
        return act, q, gone, n_create, None

    @staticmethod
    def _state_key(act):
        return [(b["id"], b["start"], b["last_active"], b["cb"], np.float32(b["rel"]).tobytes(),
                 np.float32(b["basec"]).tobytes(), b["dl"], b["lah"], b["tl"]) for b in act]

    # ------------------------------------------------------------------ one chunk
    def chunk(self, mag):
        F, N, thr = mag.shape[0], self.N, self.thr
        if not self.primed:
            self._prime(mag)
            self.index += F * N
            return
        self.stats["chunks"] += 1
        index0 = self.index
        r = self.base.copy()
        pos = r > 0
        inf = np.float32(np.inf)
        thi = np.where(pos, (thr * (r * GUARD_HI)) * np.float32(1.0001), inf).astype(np.float32)
        tlo = np.where(pos, (thr * (r * GUARD_LO)) * np.float32(0.9999), inf).astype(np.float32)
        X = mag > thi
        XU = ~(mag < tlo)
        a, b = r * GUARD_LO, r * GUARD_HI
        glo, ghi = np.minimum(a, b).astype(np.float32), np.maximum(a, b).astype(np.float32)
        rowany = XU.any(axis=1)
        # carried-in bursts: deadlines in frames of this chunk, real ids as (0, id, 0)
        carried = []
        for bst in self.act:
            d = bst["last_active"] + self.post - index0
            t = bst["start"] + self.max_len - index0
            carried.append(dict(id=(0, bst["id"], 0), start=bst["start"], last_active=bst["last_active"], cb=bst["cb"],
                                rel=bst["rel"], basec=bst["basec"], dl=0 if d <= 0 else -(-d // N), lah=None,
                                tl=BIG if self.max_len <= 0 else (-1 if t < 0 else t // N)))
        cuts = list(range(0, F, self.SEG)) + [F]
        S = len(cuts) - 1
        starts = [carried] + [[] for _ in range(S - 1)]
        Q = np.zeros(F, np.uint8)
        prev_out = None
        for rnd in range(self.max_rounds):
            self.stats["rounds"] += 1
            ver, snaps, base_final, qlist, (bmin, bmax) = self._base_pass(mag, Q, rowany, None)
            # a baseline outside the band its bin's bitmaps were made for: widen the band to what the baseline really
            # did (with a margin) and rebuild the bitmaps before anybody walks them (k_seg_base + k_seg_reclass)
            out_of_band = ~((bmin >= glo) & (bmax <= ghi))
            hopeless = out_of_band & ~((bmin > 0) & np.isfinite(bmax) & np.isfinite(base_final))
            bad = bool(hopeless.any())
            fix = out_of_band & ~hopeless
            reclass = bool(fix.any())
            if reclass:
                self.stats["rebuilds"] = self.stats.get("rebuilds", 0) + 1
                glo = np.where(fix, np.minimum(glo, bmin * np.float32(0.9)), glo).astype(np.float32)
                ghi = np.where(fix, np.maximum(ghi, bmax * np.float32(1.1)), ghi).astype(np.float32)
                thi = np.where(glo > 0, (thr * ghi) * np.float32(1.0001), inf).astype(np.float32)
                tlo = np.where(ghi > 0, (thr * glo) * np.float32(0.9999), inf).astype(np.float32)
                X = mag > thi
                XU = ~(mag < tlo)
                rowany = rowany | XU.any(axis=1)
                ver, snaps, base_final, qlist, _ = self._base_pass(mag, Q, rowany, None)   # (model only: slots for the new bits)
            ends, Qn, gones, ncs, bails = [], np.zeros(F, np.uint8), [], [], []
            for s in range(S):
                e, q, g, nc, bl = self._walk(s, cuts[s], cuts[s + 1], starts[s], mag, X, XU, ver, snaps, index0, self.sq)
                bails.append(bl)
                if bl is None:
                    ends.append(e); Qn[cuts[s]:cuts[s + 1]] = q; gones.append(g); ncs.append(nc)
                else:                                   # keep last round's view of this segment
                    ends.append(starts[s + 1] if s + 1 < S else []); gones.append([]); ncs.append(0)
                    Qn[cuts[s]:cuts[s + 1]] = Q[cuts[s]:cuts[s + 1]]
            out = (Qn.tobytes(), [self._state_key(e) for e in ends], ncs)
            same = prev_out is not None and out == prev_out and not reclass
            prev_out = out
            if same:
                self.stats["max_rounds_seen"] = max(self.stats["max_rounds_seen"], rnd + 1)
                if bad:
                    raise Bail("guard band")
                for bl in bails:
                    if bl is not None:
                        raise Bail(bl)
                break
            Q = Qn
            for s in range(1, S):
                starts[s] = ends[s - 1]
        else:
            raise Bail("no fixed point")
        # ---- commit: ids from the prefix sum of creation counts, gone list in segment order
        pref = np.concatenate(([0], np.cumsum(ncs))).astype(np.int64)

        def real_id(code):
            return code[1] if code[0] == 0 else self.next_id + 10 * int(pref[code[1]] + code[2])
        for g in gones:
            for rec in g:
                self.gone.append((real_id(rec[0]),) + rec[1:])
        self.act = []
        for bst in ends[-1]:
            la = bst["last_active"] if bst["lah"] is None else index0 + bst["lah"] * N
            self.act.append(dict(id=real_id(bst["id"]), start=bst["start"], last_active=la, cb=bst["cb"], rel=bst["rel"],
                                 basec=bst["basec"]))
        self.next_id += 10 * int(pref[-1])
        self.sq = max(self.sq - F, 0)
        nq = len(qlist)
        for q in range(max(0, nq - self.H), nq):
            self.hist[(self.hist_idx + q) % self.H] = mag[qlist[q]]
        self.hist_idx = (self.hist_idx + nq) % self.H
        self.base = base_final
        self.index = index0 + F * N

    def run(self, mag, chunk_frames=4096):
        F = mag.shape[0]
        a = 0
        while a < F:
            b = min(F, self.H) if a < self.H else min(F, a + chunk_frames)
            self.chunk(mag[a:b])
            a = b
        return self.gone
