"""Checks the static schedule of the K2s v4 kernel (dtw_stream4_kernel.cu) on the CPU: ring-slot safety of the
template row pairs, visibility before first use, window-block staging slots, exchange slots, warp occupancy.
Not product code; run `python tools/sim_stream4_schedule.py`."""
import itertools
import sys

CB, NW, SLOTS, XS = 8, 4, 16, 4


def check(m, n, band, verbose=False):
    w = max(band, abs(m - n))
    assert 3 <= w <= 20
    n_blocks = (n + CB - 1) // CB
    half = m // 2                       # row pairs that contain a needed row (rows 1 .. m-1)
    fin0 = 4 + (w + 1) // 2
    sigma = max(1, -(-(4 + w) // 4) - 4)
    P = 4 + sigma
    steps = half + sigma * (n_blocks - 1)
    kmax = (m + 1) // 2

    def u_first(B): return max(1, 4 * B + 1 - w // 2)
    def u_last(B): return 4 * B + fin0
    def S_sw(B): return u_last(B - NW) + sigma * (B - NW)   # step after which block B replaces block B-4

    # per-warp block timeline must not overlap
    for B in range(n_blocks - NW):
        end = u_last(B) + sigma * B
        nxt = u_first(B + NW) + sigma * (B + NW)
        assert nxt > end, (m, n, w, B, end, nxt)

    # --- template ring: simulate the fetch rule
    def B_min(k): return max(0, -(-(k - fin0) // 4))
    def B_max(k): return min(n_blocks - 1, (k - 1 + w // 2) // 4)
    def st_first(k): return k + sigma * B_min(k)
    def st_last(k): return k + sigma * B_max(k)
    KPRO = 4
    vis = {k: 0 for k in range(1, KPRO + 1)}   # visible from step (prologue)
    sts_step = {k: 0 for k in range(1, KPRO + 1)}
    kf = KPRO + 1
    for st in range(1, steps + 1):
        if kf <= kmax and st >= st_first(kf) - 4:
            sts_step[kf] = st + 1       # normalised + stored at the top of the next step
            vis[kf] = st + 2            # visible after that step's barrier
            kf += 1
    # reads: block B at step st reads row pair u (row 2u) at H1 and u+1 (row 2u+1) at H2
    for B in range(n_blocks):
        for st in range(1, steps + 1):
            u = st - sigma * B
            if u_first(B) <= u <= u_last(B):
                for k in (u, u + 1):
                    used = k == u or (k <= u_last(B))
                    if k > kmax or not used:
                        continue
                    if k > half + 1:
                        continue
                    assert k in vis and vis[k] <= st, ("row pair not visible", m, n, w, B, st, k, vis.get(k))
                    # slot still holds k: no later row pair stored over it yet
                    k2 = k + SLOTS
                    if k2 in sts_step:
                        assert sts_step[k2] > st, ("slot overwritten", m, n, w, B, st, k, k2, sts_step[k2])
    # --- window block staging: two slots per pair
    due = []
    for B in range(NW, n_blocks):
        for j in range(4):
            due.append((S_sw(B) - 6 + j, B, j))
    qi = 0
    col_sts = {}
    for st in range(1, steps + 1):
        if qi < len(due) and st >= due[qi][0]:
            _, B, j = due[qi]
            col_sts[(B, j)] = st + 1
            qi += 1
    for B in range(NW, n_blocks):
        sw = S_sw(B)
        if sw > steps:
            continue
        for j in range(4):
            assert (B, j) in col_sts and col_sts[(B, j)] <= sw - 1 + 0, ("block not staged", m, n, w, B, j, col_sts.get((B, j)), sw)
        if B - 2 >= NW:
            assert col_sts[(B, 0)] > S_sw(B - 2), ("stage slot overwritten", m, n, w, B)
    # --- exchange slots: writer block B at step s (row pair u) -> reader block B+1 at step s + sigma, slot u & 3
    assert sigma < XS
    busy = [0] * NW
    for B in range(n_blocks):
        for st in range(1, steps + 1):
            u = st - sigma * B
            if u_first(B) <= u <= u_last(B):
                busy[B % NW] += 1
    if verbose:
        print(dict(m=m, n=n, w=w, sigma=sigma, steps=steps, n_blocks=n_blocks, busy=busy, eff=sum(busy) / (NW * steps)))
    return True


if __name__ == "__main__":
    check(120, 100, 5, verbose=True)
    cnt = 0
    for m, n, band in itertools.product(range(2, 200, 3), range(1, 200, 3), (3, 5, 8, 12, 16, 17, 20)):
        w = max(band, abs(m - n))
        if 3 <= w <= 20:
            check(m, n, band)
            cnt += 1
    print("ok", cnt)
