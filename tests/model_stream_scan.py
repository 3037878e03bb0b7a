"""Executable model (numpy, CPU) of the streaming detector state machine of
iridium-sniffer_b200/csrc/k_detect_stream.cu -- the ALGORITHM, not the kernel: guard-banded bitmaps per
launch, deadlines kept in frames, bursts without a list (order = ascending id), events only where the
bitmaps do not prove the outcome, bail conditions.  tests/test_stream_scan_model.py runs it on the
magnitude frames of the CPU oracle and demands the oracle's burst list field for field, which pins the
arithmetic claims of DESIGN.md section 4a without a GPU.  (The CUDA kernel itself is compared with the
oracle by the -m gpu tests.)"""
import numpy as np

GUARD_LO = np.float32(0.65)
GUARD_HI = np.float32(1.5)
MAX_LANES = 32


class Bail(Exception):
    def __init__(self, reason, frame):
        super().__init__(f"bail: {reason} at frame {frame}")
        self.reason, self.frame = reason, frame


class StreamScanModel:
    def __init__(self, N, thr, half_bw, pre_len, post_len, max_burst_len, max_bursts, hist_size=512):
        self.N, self.thr, self.half_bw = N, np.float32(thr), half_bw
        self.pre, self.post, self.max_len, self.max_bursts, self.H = pre_len, post_len, max_burst_len, max_bursts, hist_size
        self.base = np.zeros(N, np.float32)
        self.hist = np.zeros((hist_size, N), np.float32)
        self.hist_idx, self.primed = 0, False
        self.act = []                      # dicts: id, start, last_active, cb, rel, basec
        self.next_id, self.index, self.sq = 0, 0, 0
        self.gone = []
        bins = np.arange(N)
        self.valid = (bins >= half_bw) & (bins < N - half_bw) & ~((bins >= N // 2 - 3) & (bins <= N // 2 + 3))
        self.stats = dict(events=0, exact_bins=0, launches=0)

    # -- baseline update of one quiet frame (simd_avx2.c:221-236: two roundings), with the guard check
    def _quiet(self, row, guard):
        old = self.hist[self.hist_idx] if self.primed else np.zeros(self.N, np.float32)
        t = self.base - old
        self.base = t + row
        self.hist[self.hist_idx] = row
        self.hist_idx += 1
        if self.hist_idx == self.H:
            self.primed, self.hist_idx = True, 0
        if guard is not None:
            lo, hi = guard
            if not np.all((self.base >= lo) & (self.base <= hi)):
                raise Bail("guard band", -1)

    def _free_mask(self):
        free = np.ones(self.N, bool)
        for b in self.act:
            free[max(b["cb"] - self.half_bw, 0):min(b["cb"] + self.half_bw, self.N - 1) + 1] = False
        return free

    def launch(self, mag):
        """frames mag[F][N] (float32); raises Bail exactly where the kernel would give up"""
        F, N, thr = mag.shape[0], self.N, self.thr
        self.stats["launches"] += 1
        index0 = self.index
        if not self.primed:
            # priming launch: every frame is quiet (burst_detect.c:426-428); no bitmaps
            for f in range(F):
                self._quiet(mag[f], None)
                if self.primed and f + 1 < F:
                    raise Bail("primed in mid-launch", f)
            self.index = index0 + F * N
            return
        if len(self.act) > MAX_LANES:
            raise Bail("more bursts than lanes", 0)
        # ---- bitmaps against the reference baseline, guard band
        r = self.base.copy()
        pos = r > 0
        inf = np.float32(np.inf)
        thi = np.where(pos, (thr * (r * GUARD_HI)) * np.float32(1.0001), inf).astype(np.float32)
        tlo = np.where(pos, (thr * (r * GUARD_LO)) * np.float32(0.9999), inf).astype(np.float32)
        X = mag > thi
        XU = ~(mag < tlo)
        a, b = r * GUARD_LO, r * GUARD_HI
        guard = (np.minimum(a, b), np.maximum(a, b))
        # ---- frame units
        PF = -(-self.post // N)
        PF0 = max(1, -(-(self.post - self.pre) // N))
        TLF = (self.max_len - self.pre) // N if self.max_len >= self.pre else -1
        for bst in self.act:
            d = bst["last_active"] + self.post - index0
            bst["dl"] = 0 if d <= 0 else -(-d // N)
            bst["lah"] = None
            t = bst["start"] + self.max_len - index0
            bst["tl"] = -1 if t < 0 else t // N
        free = self._free_mask()
        for f in range(F):
            idx = index0 + f * N
            fv = free & self.valid
            cand = bool(np.any(XU[f] & fv))
            ev = cand
            for bst in self.act:
                cb = bst["cb"]
                bst["x3"] = bool(X[f, cb - 1:cb + 2].any())
                bst["u3"] = bool(XU[f, cb - 1:cb + 2].any())
                if (not bst["x3"] and (bst["u3"] or f >= bst["dl"])) or f > bst["tl"]:
                    ev = True
            if not ev:
                for bst in self.act:
                    if bst["x3"]:
                        bst["dl"], bst["lah"] = f + PF, f
                self.sq = max(self.sq - 1, 0)
                if not self.act:
                    self._quiet(mag[f], guard)
                continue
            # ================= event frame: the reference's steps, exactly
            self.stats["events"] += 1
            if any(f > bst["tl"] for bst in self.act):
                raise Bail("too long", f)
            row = mag[f]
            for bst in self.act:                             # update_bursts (:458-469)
                hit = bst["x3"]
                if not hit and bst["u3"]:
                    for bn in range(bst["cb"] - 1, bst["cb"] + 2):
                        if XU[f, bn] and self.base[bn] > 0 and np.float32(row[bn] / self.base[bn]) > thr:
                            hit = True
                if hit:
                    bst["dl"], bst["lah"] = f + PF, f
                bst["done"] = (not hit) and f >= bst["dl"]
            # peaks: exact crossings & mask of the previous frame & search range (:522-548)
            peaks = []
            for bn in np.nonzero(XU[f] & fv)[0]:
                self.stats["exact_bins"] += 1
                if self.base[bn] > 0:
                    rel = np.float32(row[bn] / self.base[bn])
                    if rel > thr:
                        peaks.append((rel, int(bn)))
            # delete_gone_bursts (:490-518): list order = creation order = ascending id
            for bst in sorted([b for b in self.act if b["done"]], key=lambda b: b["id"]):
                la = bst["last_active"] if bst["lah"] is None else index0 + bst["lah"] * N
                self.gone.append((bst["id"], bst["start"], idx, la, bst["cb"], bst["rel"], bst["basec"]))
            if any(b["done"] for b in self.act):
                self.act = [b for b in self.act if not b["done"]]
                free = self._free_mask()                     # update_burst_mask (:482-486)
            # create_new_bursts (:556-591): strongest first, ties by bin
            peaks.sort(key=lambda p: (-p[0], p[1]))
            for rel, bn in peaks:
                if not free[bn]:
                    continue
                if len(self.act) >= MAX_LANES:
                    raise Bail("a 33rd concurrent burst", f)
                start = idx - self.pre
                self.act.append(dict(id=self.next_id, start=start, last_active=start, cb=bn, rel=rel,
                                     basec=self.base[bn], dl=f + PF0, lah=None, tl=f + TLF))
                self.next_id += 10
                free[max(bn - self.half_bw, 0):min(bn + self.half_bw, N - 1) + 1] = False
            if self.max_bursts > 0 and len(self.act) > self.max_bursts:
                raise Bail("squelch", f)
            if self.sq > 0:
                self.sq -= 1
            if not self.act:
                self._quiet(mag[f], guard)
        # ---- hand the state to the next launch: last_active back in samples
        for bst in self.act:
            if bst["lah"] is not None:
                bst["last_active"] = index0 + bst["lah"] * N
        self.index = index0 + F * N

    def run(self, mag, launch_frames=700):
        """whole recording: priming launch, then launches of launch_frames frames"""
        F = mag.shape[0]
        a = 0
        while a < F:
            b = min(F, self.H) if a < self.H else min(F, a + launch_frames)
            self.launch(mag[a:b])
            a = b
        return self.gone
