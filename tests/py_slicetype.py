"""A SECOND, independently written implementation of the lookahead's decision layer:
[x264] encoder/lookahead.c (queueing) and encoder/slicetype.c (x264_slicetype_decide, x264_slicetype_analyse,
scenecut, slicetype_path, slicetype_path_cost, the keyframe-interval rules and the coded-order shuffle), in
Python, from upstream's published behaviour -- not a transliteration of oracle/lookahead_oracle.c.

Test infrastructure.  It owns NO cost arithmetic: every slicetype_frame_cost(p0, p1, b) and every mb-tree walk is
delegated to a cost engine (the C checker opened with the maximum lookahead, so that the checker's own decision
code never runs there).  What it checks is therefore exactly the part where the C checker (and the product, whose
la_host.cu follows the same control flow) could share a misreading of upstream: which evaluations are asked for,
in which order, how their values are compared, and which frame type comes out.

    eng = CostEngine(params)                 # C checker as a cost table
    la  = PyLookahead(params, eng)
    for planes in clip: la.put(planes)
    la.flush(); la.decisions -> [dict(i_frame, i_type, b_keyframe, i_bframes, i_cost_est, ..., qp_offset)]
"""
import ctypes as C

import numpy as np

import oracle_lib as ol

AUTO, IDR, I, P, BREF, B, KEYFRAME = 0, 1, 2, 3, 4, 5, 6
COST_MAX64 = 1 << 60
LOOKAHEAD_MAX, BFRAME_MAX = 250, 16


def is_b(t):
    return t in (B, BREF)


def is_i(t):
    return t in (I, IDR, KEYFRAME)


def auto_or_i(t):
    return t == AUTO or is_i(t)


def auto_or_b(t):
    return t == AUTO or is_b(t)


class CostEngine:
    """The C checker used as a memoising cost table over display indices.  Opened with rc-lookahead = the maximum,
    so its own slicetype_decide never fires for clips shorter than that; frames are only prepared (AQ, lowres)."""

    def __init__(self, params):
        p = ol.LaParams()
        C.memmove(C.byref(p), C.byref(params), C.sizeof(p))
        p.rc_lookahead = LOOKAHEAD_MAX
        self.orc = ol.OracleLookahead(p)
        self.n = 0

    def put(self, planes):
        self.orc.put_i420(planes)
        assert not self.orc.decisions(), "the cost engine must not decide anything itself"
        self.n += 1

    def frame_cost(self, p0, p1, b):
        v = self.orc.frame_cost(p0, p1, b)
        assert v >= 0, (p0, p1, b)
        return v

    def cost_est(self, f, d0, d1, aq=False):
        return self.orc.cost_est(f, d0, d1, aq)

    def intra_mbs(self, f, d0):
        return self.orc.intra_mbs(f, d0)

    def mbtree(self, idx, types, b_intra):
        self.orc.mbtree(idx, types, b_intra)

    def qp_offset(self, f, aq=False):
        return self.orc.qp_offset(f, aq)

    def close(self):
        self.orc.close()


class Fr:
    __slots__ = ("n", "type", "forced", "scenecut_ok", "keyframe", "bframes", "rc")

    def __init__(self, n):
        self.n = n
        self.type = self.forced = AUTO
        self.scenecut_ok = True          # x264_frame_t.b_scenecut: may still be judged a scene cut
        self.keyframe = False
        self.bframes = 0
        self.rc = None                   # (d0, d1) of the evaluation rate control will read


class PyLookahead:
    def __init__(self, params, engine):
        self.p = params
        self.eng = engine
        self.next = []                   # h->lookahead->next
        self.last_nonb = None
        self.length = max(params.bframes, params.rc_lookahead)      # i_slicetype_length
        self.last_keyframe = -params.keyint_max
        self.decisions = []
        self.n_in = 0
        self.keyint_min = params.keyint_min
        if self.keyint_min <= 0:
            self.keyint_min = min(params.keyint_max // 10, params.fps_num // max(1, params.fps_den))

    # ---- queueing ([x264] x264_lookahead_put_frame / lookahead_slicetype_decide) ------------------------------
    def put(self, planes):
        self.eng.put(planes)
        self.next.append(Fr(self.n_in))
        self.n_in += 1
        while len(self.next) > self.length:
            self._decide_and_shift()

    def flush(self):
        while self.next:
            self._decide_and_shift()

    # ---- cost helpers: frames[] are Fr records, the engine speaks display indices ------------------------------
    def _fc(self, frames, p0, p1, b):
        return self.eng.frame_cost(frames[p0].n, frames[p1].n, frames[b].n)

    def _path_cost(self, frames, path, threshold):
        """Sum of the frame costs a type string implies ('P'/'I' = non-B, 'B'), stopping early above `threshold`."""
        pyramid = self.p.b_pyramid
        cost, cur, k = 0, 0, 0           # path[k] is the type of frames[k + 1]
        while k < len(path):
            nxt = k
            while path[nxt] == "B":
                nxt += 1
            nb = nxt + 1                 # index in frames[] of the next non-B
            cost += self._fc(frames, cur, nb, nb) if path[nxt] == "P" else self._fc(frames, nb, nb, nb)
            if cost > threshold:
                break
            first_b = k + 1
            if pyramid and nb - cur > 2:
                mid = cur + (nb - cur) // 2
                cost += self._fc(frames, cur, nb, mid)
                for b in range(first_b, mid):
                    if cost >= threshold:
                        break
                    cost += self._fc(frames, cur, mid, b)
                for b in range(mid + 1, nb):
                    if cost >= threshold:
                        break
                    cost += self._fc(frames, mid, nb, b)
            else:
                for b in range(first_b, nb):
                    if cost >= threshold:
                        break
                    cost += self._fc(frames, cur, nb, b)
            k, cur = nxt + 1, nb
        return cost

    def _best_path(self, frames, length, best):
        """b-adapt 2: extend the best paths of shorter prefixes by 'B'*k + 'P' and keep the cheapest admissible one."""
        best_cost, best_possible, keep = COST_MAX64, False, None
        for k in range(min(self.p.bframes + 1, length)):
            ln = length - (k + 1)
            path = list(best[ln % (BFRAME_MAX + 1)][:ln].ljust(ln, "\0")) + ["B"] * k + ["P"]
            possible = True
            for i in range(1, length + 1):
                t = frames[i].type
                if t == AUTO:
                    continue
                if is_b(t):
                    possible = possible and (i < ln or i == length or path[i - 1] == "B")
                else:
                    possible = possible and (i < ln or path[i - 1] != "B")
                    path[i - 1] = "I" if is_i(t) else "P"
            if possible or not best_possible:
                if possible and not best_possible:
                    best_cost = COST_MAX64
                cost = self._path_cost(frames, "".join(path), best_cost)
                if cost < best_cost:
                    best_cost, best_possible, keep = cost, possible, "".join(path)
        if keep is None:     # upstream copies the untouched scratch buffer in that case; cannot happen with finite costs
            keep = "".join(path)
        best[length % (BFRAME_MAX + 1)] = keep

    # ---- scene cuts ----------------------------------------------------------------------------------------------
    def _is_cut(self, frames, p0, p1):
        f = frames[p1]
        self._fc(frames, p0, p1, p1)
        icost = self.eng.cost_est(f.n, 0, 0)
        pcost = self.eng.cost_est(f.n, p1 - p0, 0)
        gop = f.n - self.last_keyframe
        tmax = np.float32(self.p.scenecut / 100.0)
        tmin = np.float32(tmax * 0.25)
        kmin, kmax = self.keyint_min, self.p.keyint_max
        if kmin == kmax:
            tmin = tmax
        # upstream computes the bias in single precision
        if gop <= kmin // 4:
            bias = np.float32(tmin / np.float32(4))
        elif gop <= kmin:
            bias = np.float32(np.float32(tmin * np.float32(gop)) / np.float32(kmin))
        else:
            bias = np.float32(tmin + np.float32(np.float32((tmax - tmin) * np.float32(gop - kmin)) / np.float32(kmax - kmin)))
        return pcost >= (1.0 - float(bias)) * icost

    def _scenecut(self, frames, p0, p1, real, num_frames, max_search):
        if real and self.p.bframes:
            # flash rejection: a frame is no cut if the picture after the flash still predicts well from before it
            far = p0 + 1 + (self.p.bframes if self.p.b_adapt == 2 else 1)
            maxp1 = min(far, num_frames)
            for c in range(p1, maxp1 + 1):
                if not self._is_cut(frames, p0, c):
                    for i in range(c, p0, -1):
                        frames[i].scenecut_ok = False
            for c in range(p0, maxp1 + 1):
                if far > max_search or (c < maxp1 and self._is_cut(frames, c, maxp1)):
                    frames[c].scenecut_ok = False
        if not frames[p1].scenecut_ok:
            return False
        return self._is_cut(frames, p0, p1)

    # ---- [x264] x264_slicetype_analyse ------------------------------------------------------------------------
    def _analyse(self, intra_minigop):
        p = self.p
        max_search = min(len(self.next), LOOKAHEAD_MAX, self.length + 1 - intra_minigop)     # b_deterministic
        keyframe = bool(intra_minigop)
        if self.last_nonb is None:
            return
        frames = [self.last_nonb] + self.next[:max_search]
        count = len(frames) - 1
        if count == 0:
            if p.b_mbtree:
                self._mbtree(frames, 0, keyframe)
            return
        keyint_limit = p.keyint_max - frames[0].n + self.last_keyframe - 1
        orig_num = num = min(count, keyint_limit)
        if p.b_psy and p.b_mbtree:
            num = count
        elif p.open_gop and num < count:
            num += 1
        elif num == 0:
            frames[1].type = I
            return

        if auto_or_i(frames[1].type) and p.scenecut and self._scenecut(frames, 0, 1, True, orig_num, max_search):
            if frames[1].type == AUTO:
                frames[1].type = I
            return

        for j in range(1, num + 1):
            if frames[j].type == KEYFRAME:
                frames[j].type = I if p.open_gop else IDR
        for j in range(2, num + 1):
            if frames[j].type == IDR and auto_or_b(frames[j - 1].type):
                frames[j - 1].type = P

        analysed = num
        if p.bframes:
            if p.b_adapt == 2:
                if num > 1:
                    best = [""] * (BFRAME_MAX + 1)
                    best[1] = "P"
                    for j in range(2, num + 1):
                        self._best_path(frames, j, best)
                    chosen = best[num % (BFRAME_MAX + 1)]
                    for j in range(1, num):
                        if chosen[j - 1] != "B":
                            if auto_or_b(frames[j].type):
                                frames[j].type = P
                        elif frames[j].type == AUTO:
                            frames[j].type = B
            elif p.b_adapt == 1:
                anchor, left = 0, p.bframes
                for j in range(1, num):
                    if j - 1 > 0 and is_b(frames[j - 1].type):
                        left -= 1
                    else:
                        anchor, left = j - 1, p.bframes
                    if not left:
                        if auto_or_b(frames[j].type):
                            frames[j].type = P
                        continue
                    if frames[j].type != AUTO:
                        continue
                    if is_b(frames[j + 1].type):
                        frames[j].type = P
                        continue
                    nb = j - anchor - 1
                    sub = frames[anchor:]
                    cost_p = self._path_cost(sub, "B" * nb + "PP", COST_MAX64)
                    cost_b = self._path_cost(sub, "B" * nb + "BP", cost_p)
                    frames[j].type = B if cost_b < cost_p else P
            else:
                left = p.bframes
                for j in range(1, num):
                    if not left:
                        if auto_or_b(frames[j].type):
                            frames[j].type = P
                    elif frames[j].type == AUTO:
                        frames[j].type = P if is_b(frames[j + 1].type) else B
                    left = left - 1 if is_b(frames[j].type) else p.bframes
            if auto_or_b(frames[num].type):
                frames[num].type = P

            lead_b = 0
            while lead_b < num and is_b(frames[lead_b + 1].type):
                lead_b += 1
            # a scene cut inside the first mini-GOP ends it there
            for j in range(1, lead_b + 1):
                if frames[j].forced == AUTO and auto_or_i(frames[j + 1].forced) and p.scenecut and \
                        self._scenecut(frames, j, j + 1, False, orig_num, max_search):
                    frames[j].type = P
                    analysed = j
                    break
            reset_start = 1 if keyframe else min(lead_b + 2, analysed + 1)
        else:
            for j in range(1, num + 1):
                if auto_or_b(frames[j].type):
                    frames[j].type = P
            reset_start = 1 if keyframe else 2

        if p.b_mbtree:
            self._mbtree(frames, min(num, p.keyint_max), keyframe)

        # keyframe interval
        last_key, last_possible = self.last_keyframe, 0
        j = 1
        while j <= num:
            f = frames[j]
            dist = f.n - last_key
            if auto_or_i(f.forced) and (p.open_gop or not is_b(frames[j - 1].forced)):
                last_possible = j
            if dist >= p.keyint_max:
                if last_possible not in (0, j):
                    j = last_possible
                    f = frames[j]
                    dist = f.n - last_key
                last_possible = 0
                if f.type != IDR:
                    f.type = I if p.open_gop else IDR
            if f.type == I and dist >= self.keyint_min:
                if p.open_gop:
                    last_key = f.n
                elif f.forced != I:
                    f.type = IDR
            if f.type == IDR:
                last_key = f.n
                if j > 1 and is_b(frames[j - 1].type):
                    frames[j - 1].type = P
            j += 1

        for j in range(reset_start, num + 1):
            frames[j].type = frames[j].forced

    def _mbtree(self, frames, num_frames, b_intra):
        fr = frames[:num_frames + 1]
        self.eng.mbtree([f.n for f in fr], [f.type for f in fr], int(b_intra))

    # ---- [x264] x264_slicetype_decide + the shift of lookahead_slicetype_decide ---------------------------------
    def _decide_and_shift(self):
        p = self.p
        nxt = self.next
        if (p.bframes and p.b_adapt) or p.scenecut or p.b_mbtree:
            self._analyse(0)
        visible = len(nxt)
        nb = brefs = 0
        while True:
            f = nxt[nb]
            if f.type == BREF and p.b_pyramid < 2 and brefs == p.b_pyramid:
                f.type = B
            elif f.type == BREF and p.b_pyramid == 2 and brefs and p.frame_reference <= brefs + 3:
                f.type = B
            if f.type == KEYFRAME:
                f.type = I if p.open_gop else IDR
            if f.n - self.last_keyframe >= p.keyint_max:
                forced = I if (p.open_gop and self.last_keyframe >= 0) else IDR
                if f.type in (AUTO, I):
                    f.type = forced
                if f.type != IDR and not (p.open_gop and f.type == I):
                    f.type = forced
            if f.type == I and f.n - self.last_keyframe >= self.keyint_min:
                if p.open_gop:
                    self.last_keyframe = f.n
                    f.keyframe = True
                else:
                    f.type = IDR
            if f.type == IDR:
                self.last_keyframe = f.n
                f.keyframe = True
                if nb > 0:
                    nb -= 1
                    nxt[nb].type = P
            if nb == p.bframes or nb + 1 >= visible:
                if f.type == AUTO or is_b(f.type):
                    f.type = P
            if f.type == BREF:
                brefs += 1
            if f.type == AUTO:
                f.type = B
            elif not is_b(f.type):
                break
            nb += 1
        nxt[nb].bframes = nb
        if p.b_pyramid and nb > 1 and not brefs:
            nxt[(nb - 1) // 2].type = BREF
            brefs += 1

        # the cost rate control will read for the frame that closes the mini-GOP
        frames = [self.last_nonb] + nxt[:nb + 1]
        b = p1 = nb + 1
        p0 = b if is_i(nxt[nb].type) else 0
        self.eng.frame_cost(frames[p0].n, frames[p1].n, frames[b].n)
        nxt[nb].rc = (b - p0, p1 - b)

        coded = [nxt[nb]] + [f for f in nxt[:nb] if f.type == BREF] + [f for f in nxt[:nb] if f.type != BREF]
        self.last_nonb = nxt[nb]
        shift = nb + 1
        del nxt[:shift]
        if p.b_mbtree and is_i(self.last_nonb.type):
            self._analyse(shift)
        for f in coded:
            d = dict(i_frame=f.n, i_type=f.type, b_keyframe=int(f.keyframe), i_bframes=f.bframes,
                     i_cost_est=-1, i_cost_est_aq=-1, i_intra_mbs=-1,
                     qp_offset=self.eng.qp_offset(f.n), qp_offset_aq=self.eng.qp_offset(f.n, True))
            if f.rc is not None:
                d["i_cost_est"] = self.eng.cost_est(f.n, f.rc[0], f.rc[1])
                d["i_cost_est_aq"] = self.eng.cost_est(f.n, f.rc[0], f.rc[1], True)
                d["i_intra_mbs"] = self.eng.intra_mbs(f.n, f.rc[0])
            self.decisions.append(d)
