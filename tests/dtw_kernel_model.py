"""
Lane-level numpy model of the CUDA wavefront DTW kernel
(whisper-timestamped_b200/csrc/dtw.cu: dtw_fill_strip + the row-wise backtrack).

It mirrors the kernel's data flow step by step — 31-row strips, lane 0 as the row above the
strip, shfl_up for `up`, last step's `up` as `diag`, the skewed 64-slot shared-memory staging
ring (row k of tile t+1 is issued at step k of tile t, so a slot is rewritten only after its last read), 2-bit direction fields packed by step index, boundary row
hand-over between strips, and the clz-based one-step-per-token backtrack — so that the index
arithmetic of the kernel can be checked on the CPU (no GPU in the build container) against the
oracle.  It is a test helper, not product code.
"""
import numpy as np

RS = 31
RING = 64
PITCH = RING + 1


def niter_of(F):
    return (F + RS - 1 + 31) // 32


def model_dtw(cost):
    """cost: [T, F] float32/float64.  Returns jumps (T+1 int32) exactly as the kernel computes."""
    cost = np.asarray(cost)
    T, F = cost.shape
    lanes = np.arange(32)
    niter = niter_of(F)
    W = 2 * niter
    nstrips = (T + RS - 1) // RS
    dirs = np.zeros((nstrips, W, 32), dtype=np.uint32)
    bnd = np.full(F + 4, np.nan)
    tile = np.zeros((32, PITCH), dtype=cost.dtype)
    INF = np.inf
    flat = cost.reshape(-1)

    for strip in range(nstrips):
        row0 = strip * RS
        Ts = min(RS, T - row0)
        first = strip == 0
        write_bnd = strip + 1 < nstrips
        cur = np.full(32, INF)
        upprev = np.full(32, INF)
        if first:
            upprev[1] = 0.0
        else:
            cur[0] = bnd[0]
        acc = np.zeros(32, dtype=np.uint32)
        base = row0 * F

        def issue_tile(u, rows):
            # cp.async of tile u (columns 32u..32u+31), strip rows in `rows` (tile-row k <-> strip row k-1)
            col = np.minimum(32 * u + lanes, F - 1)
            for k in rows:
                tile[k, (lanes + 32 * u + k - 1) & (RING - 1)] = flat[base + (k - 1) * F + col]

        issue_tile(0, range(1, Ts + 1))
        bndnext = np.full(32, INF)
        if not first:
            idx = 1 + lanes
            bndnext = np.where(idx < F, bnd[np.minimum(idx, F - 1)], INF)

        for t in range(niter):
            if not first:
                bndreg = bndnext
                idx = 32 * (t + 1) + 1 + lanes
                bndnext = np.where(idx < F, bnd[np.minimum(idx, F + 3)], INF)
            cb = (32 * t) & (RING - 1)
            for k in range(32):
                if 1 <= k <= Ts:
                    issue_tile(t + 1, [k])      # row k of the NEXT tile is issued at step k (cp.async, one row per step)
                s = 32 * t + k
                l = tile[lanes, cb + k].astype(np.float64)
                # staging check: every active cell must see its own cost value
                for L in range(1, Ts + 1):
                    j = s + 1 - L
                    if 0 <= j < F:
                        assert l[L] == np.float64(cost[row0 + L - 1, j]), (strip, t, k, L, j)
                up = np.empty(32)
                up[1:] = cur[:-1]
                up[0] = cur[0]
                diag = upprev.copy()
                upprev = up.copy()
                c1, c2, c3 = diag + l, cur + l, up + l
                p2 = c2 < c1
                m = np.where(p2, c2, c1)
                p3 = c3 < m
                best = np.where(p3, c3, m)
                cur = best.copy()
                if (k & 15) == 0:
                    acc[:] = 0
                acc |= (p2.astype(np.uint32) << np.uint32(2 * (k & 15)))
                acc |= (p3.astype(np.uint32) << np.uint32(2 * (k & 15) + 1))
                if (k & 15) == 15:
                    dirs[strip, 2 * t + (k >> 4), :] = acc
                if write_bnd:
                    j = s - (Ts - 1)
                    if 0 <= j < F:
                        bnd[j] = best[Ts]
                if not first:
                    cur[0] = bndreg[k]

    def clz(x):
        return 32 - int(x).bit_length()

    jumps = np.zeros(T + 1, dtype=np.int32)
    jumps[T] = F - 1
    i, j = T - 1, F - 1
    while i > 0:
        strip, ln = i // RS, i % RS + 1
        s = j + ln - 1
        w, pos = s >> 4, s & 15
        while True:
            x = int(dirs[strip, w, ln])
            lo, hi = x & 0x55555555, (x >> 1) & 0x55555555
            nl = 0x55555555 & ~(lo & ~hi)
            m = nl & (0xFFFFFFFF >> (30 - 2 * pos))
            if m:
                kf = (31 - clz(m)) >> 1
                break
            if w == 0:
                kf = 0
                break
            w -= 1
            pos = 15
        jj = max(w * 16 + kf - (ln - 1), 0)
        is_up = (x >> (2 * kf + 1)) & 1
        jumps[i] = jj
        j = jj - 1 if (not is_up and jj > 0) else jj
        i -= 1
    jumps[0] = 0
    return jumps


# ----------------------------------------------------------------------------------------------------------------
# Model of the single-strip fast path (dtw.cu: dtw_small_kernel<TC, LA, .>): un-skewed ring of NT = LA + 1 + ceil(31/TC)
# tiles + mirror of slot 0, one bulk copy per row and tile issued LA tiles ahead, per-lane read base
# p_L = (TC t - L + 1) mod ring.
# `late=True` lets every bulk copy land at the last possible moment (just before the mbarrier wait of its tile),
# `late=False` at issue time: the kernel must be right for both, i.e. no slot is overwritten while still needed and
# no slot is read before its tile's wait.
def small_geometry(TC, LA):
    NT = LA + 1 + (31 + TC - 1) // TC
    ring = NT * TC
    assert ring % 32 == 0
    return NT, ring, ring + TC


def model_dtw_small(cost, late=False, TC=16, LA=1):
    SM_TC, SM_LA = TC, LA
    SM_NT, SM_RING, SM_PITCH = small_geometry(TC, LA)
    cost = np.asarray(cost, dtype=np.float32)
    T, F = cost.shape
    assert T <= RS
    P = (F + 3) & ~3
    padded = np.zeros((T, P), dtype=np.float32)
    padded[:, :F] = cost
    niter = niter_of(F)
    nit = niter * (32 // SM_TC)
    ntile = (P + SM_TC - 1) // SM_TC
    lanes = np.arange(32)
    tile = np.zeros((32, SM_PITCH), dtype=np.float32)
    dirs = np.zeros((2 * niter, 32), dtype=np.uint32)
    pending = {}
    acc = np.zeros(32, dtype=np.uint32)

    def issue(u):
        ncol = min(SM_TC, P - SM_TC * u)
        slot = u % SM_NT
        writes = []
        for L in range(1, T + 1):
            data = padded[L - 1, SM_TC * u: SM_TC * u + ncol]
            writes.append((L, slot * SM_TC, data))
            if slot == 0:
                writes.append((L, SM_RING, data))
        pending[u] = writes

    def land(u):
        for (L, off, data) in pending.pop(u, []):
            tile[L, off: off + len(data)] = data

    cur = np.full(32, np.inf)
    upprev = np.full(32, np.inf)
    upprev[1] = 0.0
    p = (SM_RING - lanes + 1) % SM_RING
    for u in range(SM_LA):
        if u < ntile:
            issue(u)
            if not late:
                land(u)
    for t in range(nit):
        if t < ntile:
            land(t)                                    # mbarrier wait: tile t is complete from here on
        if t + SM_LA < ntile:
            issue(t + SM_LA)
            if not late:
                land(t + SM_LA)
        for k in range(SM_TC):
            assert np.all(p + k < SM_PITCH)
            # bank-conflict freedom of the read of this step (the point of the ring / pitch geometry)
            assert len(set(((lanes * SM_PITCH + p + k) % 32).tolist())) == 32
            l = tile[lanes, p + k].astype(np.float64)
            up = np.concatenate(([cur[0]], cur[:-1]))
            c1, c2, c3 = upprev + l, cur + l, up + l
            upprev = up
            p2 = c2 < c1
            m = np.where(p2, c2, c1)
            p3 = c3 < m
            cur = np.where(p3, c3, m)
            step = t * SM_TC + k
            if step % 16 == 0:
                acc = np.zeros(32, dtype=np.uint32)
            sh_ = np.uint32(2 * (step % 16))
            acc |= (p2.astype(np.uint32) << sh_) | (p3.astype(np.uint32) << (sh_ + np.uint32(1)))
            if step % 16 == 15:
                dirs[step // 16] = acc
        p = (p + SM_TC) % SM_RING
    # row-wise backtrack on the packed words (same arithmetic as dtw_backtrack_jumps)
    jumps = np.zeros(T + 1, dtype=np.int32)
    jumps[T] = F - 1
    i, j = T - 1, F - 1
    while i > 0:
        ln = i + 1
        s = j + ln - 1
        w, pos = s >> 4, s & 15
        while True:
            x = int(dirs[w, ln])
            lo, hi = x & 0x55555555, (x >> 1) & 0x55555555
            nonleft = 0x55555555 & ~(lo & ~hi)
            msk = nonleft & (0xffffffff >> (30 - 2 * pos))
            if msk:
                kf = (msk.bit_length() - 1) >> 1
                break
            if w == 0:
                kf = 0
                break
            w, pos = w - 1, 15
        jj = max(w * 16 + kf - (ln - 1), 0)
        is_up = (x >> (2 * kf + 1)) & 1
        jumps[i] = jj
        j = jj - 1 if (not is_up and jj > 0) else jj
        i -= 1
    return jumps


# ----------------------------------------------------------------------------------------------------------------
# Index model of lean_mma_kernel (decode_steps.cu): small-batch matrix-vector products on mma.sync.m16n8k16 with a
# PERMUTED k order, so that a thread's weight fragment for two MMAs is one 16-byte load of 8 consecutive k.
# Inside every block of 32 k, word w = 4 s + q of the staged activations (two bf16) holds actual k = 8 q + 2 s + {0, 1}.
def mma_model(x, w):
    """x [16, K] float, w [8, K] float -> D [16, 8] through the fragment definitions of mma.m16n8k16 (PTX ISA):
    A: a0 (row g, k 2t..2t+1), a1 (row g+8, same k), a2 (row g, k 8+2t..), a3 (row g+8, k 8+2t..);  B: b0 (k 2t..2t+1, n g),
    b1 (k 8+2t.., n g);  D: d0,d1 (row g, n 2t..2t+1), d2,d3 (row g+8, ...) with g = lane >> 2, t = lane & 3."""
    x = np.asarray(x, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    K = x.shape[1]
    assert K % 32 == 0
    # staging: permuted activation words, xs_perm[row, kb, word, e]
    xp = np.zeros((16, K // 32, 16, 2))
    for kb in range(K // 32):
        for q in range(4):
            for s in range(4):
                for e in range(2):
                    xp[:, kb, 4 * s + q, e] = x[:, 32 * kb + 8 * q + 2 * s + e]
    D = np.zeros((16, 8))
    for kb in range(K // 32):
        # The MMA contracts over k-slots across the four threads of a group for A (rows) and across groups for B (n):
        # emulate it exactly: build the 16x16 A tile and 16x8 B tile from the fragments and multiply.
        for m, (sa, sb) in enumerate(((0, 1), (2, 3))):
            A = np.zeros((16, 16))
            B = np.zeros((16, 8))
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                A[g, 2 * t: 2 * t + 2] = xp[g, kb, 4 * sa + t]
                A[g + 8, 2 * t: 2 * t + 2] = xp[g + 8, kb, 4 * sa + t]
                A[g, 8 + 2 * t: 8 + 2 * t + 2] = xp[g, kb, 4 * sb + t]
                A[g + 8, 8 + 2 * t: 8 + 2 * t + 2] = xp[g + 8, kb, 4 * sb + t]
                words = w[g, 32 * kb + 8 * t: 32 * kb + 8 * t + 8].reshape(4, 2)
                B[2 * t: 2 * t + 2, g] = words[2 * m]
                B[8 + 2 * t: 8 + 2 * t + 2, g] = words[2 * m + 1]
            D += A @ B
    return D


def model_dtw_lane(cost, TR=24, stale=None, NC=2, G=2):
    """dtw_lane_kernel<TR, NC, G> for the G lanes of ONE matrix: band b owns rows [b RL, (b + 1) RL) in a register
    column D[RL] and runs b column groups behind band b - 1, whose last row it gets by shuffle once per group; NC columns
    advance together, skewed by one row each (chain c sits on cell (s - c, j + c) in step s and takes `left` / `diag` from
    what chain c - 1 produced one / two steps earlier); direction words [16-column group][row], flushed per 16 columns;
    clz backtrack by the first band's lane.  Cells outside the matrix (rows >= T, columns >= the padded pitch) read
    `stale` values, as the kernel reads never-written shared memory there.  cost: float32 [T, F], all <= 0 with
    cost[0, 0] < 0."""
    cost = np.asarray(cost, np.float32)
    T, F = cost.shape
    assert 1 <= T <= TR <= 32 and TR % 8 == 0 and NC in (2, 4) and G in (1, 2, 4) and G - 1 < 8 // NC
    RL, GPT = TR // G, 8 // NC
    P = (F + 3) & ~3
    rng = np.random.default_rng(1234)
    ntile = (F + 7) >> 3
    ncols = (ntile + 2) * 8
    stage = (-rng.random((TR, ncols)).astype(np.float32) * 7) if stale is None else np.full((TR, ncols), stale, np.float32)
    stage[:T, :F] = cost
    stage[:T, F:P] = 0                                   # padding columns hold zeros

    def less(a, b):                                      # the integer compare on strictly negative doubles / +inf
        ua = np.array([a], np.float64).view(np.uint64)[0]
        ub = np.array([b], np.float64).view(np.uint64)[0]
        return bool(ua > ub)

    INF = np.float64(np.inf)
    ngroups = (F + 15) >> 4
    dirs = np.zeros((ngroups + 2) * TR, np.uint32)

    def didx(g16, row):                                  # dtw_lane_dir_index: blocks of 2 column groups x 4 rows
        return (((g16 >> 1) * (TR // 4) + (row >> 2)) << 3) + ((g16 & 1) << 2) + (row & 3)

    def cell(diag, left, up, l):
        l = np.float64(l)
        with np.errstate(invalid="ignore"):
            c1, c2, c3 = diag + l, left + l, up + l
        p2 = less(c2, c1)
        m = c2 if p2 else c1
        p3 = less(c3, m)
        return (c3 if p3 else m), p2, p3

    class Lane:
        pass

    lanes = []
    for b in range(G):
        ln = Lane()
        ln.D, ln.acc, ln.last, ln.plast = [INF] * RL, [0] * RL, [INF] * NC, INF
        lanes.append(ln)

    def cols(ln, rows0, j, diag0, bnd, one, two):
        up, h1, h2 = list(bnd), [INF] * NC, [INF] * NC
        h1[0] = bnd[0]
        diagA = diag0
        for s in range(RL + NC - 1):
            out = [INF] * NC
            for c in range(NC):
                r = s - c
                if r < 0 or r >= RL:
                    out[c] = bnd[c] if r == -1 else INF
                    continue
                left = ln.D[r] if c == 0 else h1[c - 1]
                diag = diagA if c == 0 else h2[c - 1]
                cur, p2, p3 = cell(diag, left, up[c], stage[rows0 + r, j + c])
                if p2:
                    ln.acc[r] |= one << (2 * c)
                if p3:
                    ln.acc[r] |= two << (2 * c)
                up[c] = cur
                out[c] = cur
                if c == 0:
                    diagA = left
                if c == NC - 1:
                    ln.D[r] = cur
                if r == RL - 1:
                    ln.last[c] = cur
            h2, h1 = h1, out

    nsteps = ntile * GPT + (G - 1)
    for gs in range(nsteps):
        shuffled = [(list(lanes[b - 1].last), lanes[b - 1].plast) if b > 0 else None for b in range(G)]   # before anyone moves
        for b, ln in enumerate(lanes):
            mg = gs - b
            if b == 0:
                bnd, diag0 = [INF] * NC, (np.float64(0.0) if mg == 0 else INF)
            else:
                bnd, diag0 = shuffled[b]
            mgc = max(mg, 0)
            j = NC * mgc
            sh = 2 * (j & 15)
            ln.plast = ln.last[NC - 1]
            cols(ln, b * RL, j, diag0, bnd, 1 << sh, 2 << sh)
            if mg < 0:
                ln.D, ln.acc, ln.last, ln.plast = [INF] * RL, [0] * RL, [INF] * NC, INF
            if mg >= 0 and ((j + NC) & 15) == 0:
                g16 = j >> 4
                if g16 < ngroups:
                    for q in range(RL):
                        dirs[didx(g16, b * RL + q)] = ln.acc[q] & 0xffffffff
                ln.acc = [0] * RL
    for b, ln in enumerate(lanes):
        jend = NC * (nsteps - b)
        if jend & 15:
            g16 = jend >> 4
            if g16 < ngroups:
                for q in range(RL):
                    dirs[didx(g16, b * RL + q)] = ln.acc[q] & 0xffffffff

    def nonleft(x):
        lo, hi = x & 0x55555555, (x >> 1) & 0x55555555
        return 0x55555555 & ~(lo & ~hi) & 0xffffffff

    jumps = np.zeros(T + 1, np.int64)
    jumps[T] = F - 1
    i, j = T - 1, F - 1
    while i > 0:
        g, pos = j >> 4, j & 15
        while True:
            x = int(dirs[didx(g, i)])
            msk = nonleft(x) & (0xffffffff >> (30 - 2 * pos))
            if msk:
                kf = (msk.bit_length() - 1) >> 1
                break
            if g == 0:
                kf = 0
                break
            g, pos = g - 1, 15
        jj = g * 16 + kf
        is_up = (x >> (2 * kf + 1)) & 1
        jumps[i] = jj
        j = jj - 1 if (not is_up and jj > 0) else jj
        i -= 1
    jumps[0] = 0
    return jumps


def lane_ring_schedule_ok(NC, G, ntile, early, issue_step=None):
    """The two-slot staging ring of dtw_lane_kernel<TR, NC, G> as a schedule: tile t (8 columns) goes to slot t & 1; the
    step with gs % GPT == 0 first waits for every copy issued so far, the step with gs % GPT == G - 1 then issues tile
    gs // GPT + 1; band b reads column group gs - b in step gs.  `early`: a copy lands the moment it is issued (the
    earliest it can overwrite its slot) — otherwise only at the next wait (the latest it may).  `issue_step` overrides
    G - 1 (to show that the rule is tight).  Returns True when every read of every band finds its own tile in the slot."""
    GPT = 8 // NC
    assert G - 1 < GPT
    issue_step = G - 1 if issue_step is None else issue_step
    slots = {0: None, 1: None}                           # slot -> tile it holds
    pending = []                                         # issued, not yet landed

    def land():
        for t in pending:
            slots[t & 1] = t
        del pending[:]

    def issue(t):
        if t < ntile:
            pending.append(t)
            if early:
                land()

    issue(0)
    for gs in range(ntile * GPT + G - 1):
        t, p = divmod(gs, GPT)
        if p == 0 and t < ntile:
            land()                                       # cp.async.wait_group 0 + __syncwarp
        # this step's shared-memory reads happen before the copies issued in the same step (program order)
        for b in range(G):
            mg = gs - b
            if 0 <= mg < ntile * GPT:
                tm = mg // GPT
                if slots[tm & 1] != tm:
                    return False
        if p == issue_step:
            issue(t + 1)
    return True
