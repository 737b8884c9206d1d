#!/usr/bin/env python
"""Summarise a BWGR_TRACE file of the pipelined sweep: globaltimer (ns) stamps of every CTA, [cta][block][16].
worker k: 0 dE polled, 1 DL written, 2 U issued, 3 u_done seen, 4 U-epilogue done, 5 G issued (row b), 6 g_done seen (row c),
7 g-epilogue done / reds issued (row c), 12 tile gather issued, 13 tile landed.  solver (last cta): 8 h polled, 9 corrected, 10 solved, 11 published."""
import sys
import numpy as np
raw = np.fromfile(sys.argv[1], dtype=np.int64)
G, nb, K, D = raw[:4]
t = raw[4:].reshape(G, nb, K).astype(np.float64)
W = G - 1
lo, hi = 50, min(350, nb - 3)
S = t[W]
print("grid", G, "blocks", nb, "D", D)
print("solver period ns", np.median(np.diff(S[lo:hi, 8])))
print(" solver: h polled->corr %d | corr->solved %d | solved->published %d | published->next h polled %d" % (
    tuple(np.median(np.diff(S[lo:hi, 8:12], axis=1), axis=0)) + (np.median(S[lo + 1:hi + 1, 8] - S[lo:hi, 11]),)))
b = np.arange(lo, hi)
c = b + 1 + D
if K >= 32:
    print(" prefetch cycles (block nb): slot free->gram cp.async issued %d | ->marker inputs written %d | ->gram landed/raw ready %d | ->inverses done %d ; in_ready -> solve(nb) h polled %d ns" % (
        np.median(S[b, 17] - S[b, 16]), np.median(S[b, 18] - S[b, 17]), np.median(S[b, 23] - S[b, 18]), np.median(S[b, 19] - S[b, 23]), np.median(S[b, 8] - S[b, 3])))
    print(" prefetch start(nb) - solver published(nb-3): %d ns ; prefetch period %d ns" % (np.median(S[b, 0] - S[b - 3, 11]), np.median(np.diff(S[lo:hi, 0]))))
w = t[:W]
# per worker medians
def med(x):
    return np.median(x, axis=1)
pub = S[b, 11][None, :]
d_poll = med(w[:, b, 0] - pub)            # published -> polled
d_dl = med(w[:, b, 2] - w[:, b, 0])       # polled -> U issued
d_u = med(w[:, b, 4] - w[:, b, 2])        # U issued -> U-epi done
d_g = med(w[:, b, 5] - w[:, b, 4])        # U-epi done -> G issued
d_ge = med(w[:, c, 7] - w[:, b, 5])       # G issued -> reds issued
hp = S[c, 8][None, :]
d_h = med(hp - w[:, c, 7])                # reds issued -> solver polled h
for name, v in (("published->polled", d_poll), ("polled->U issued", d_dl), ("U issued->U-epi done", d_u), ("U-epi done->G issued", d_g),
                ("G issued->reds issued", d_ge), ("reds issued->h polled by solver", d_h)):
    print(" %-34s min %7.0f med %7.0f max %7.0f (worker %d)" % (name, v.min(), np.median(v), v.max(), int(v.argmax())))
if K >= 32:
    nr = min(W, 128)
    last = w[:, c, 7].max(axis=0)
    print(" last partial store -> reducer rows complete: med %d max-over-reducers med %d | row complete -> result stored %d | last result stored -> solver polled %d" % (
        np.median(w[:nr, c, 14] - last[None, :]), np.median(w[:nr, c, 14].max(axis=0) - last), np.median(w[:nr, c, 15] - w[:nr, c, 14]),
        np.median(S[c, 8] - w[:nr, c, 15].max(axis=0))))
last_red = w[:, c, 7].max(axis=0)
print(" last red of any worker -> h polled: med %d" % np.median(S[c, 8] - last_red))
print(" published -> last worker polled: med %d" % np.median(w[:, b, 0].max(axis=0) - S[b, 11]))
if K >= 32:
    for wk in (0, W // 2, W - 1):
        x = t[wk]
        ck = lambda k1, b1, k0, b0: np.median(x[b1, 16 + k1] - x[b0, 16 + k0])
        print(" worker %d clock64: polled->DL %d | DL->U issued %d | U issued->u_done %d | u_done->U-epi %d | U-epi->G issued %d | G issued->g_done %d | g_done->stores %d | stores->next polled %d" % (
            wk, ck(1, b, 0, b), ck(2, b, 1, b), ck(3, b, 2, b), ck(4, b, 3, b), ck(5, b, 4, b), ck(6, c, 5, b), ck(7, c, 6, c), ck(0, b + 1, 7, c)))
    x = t[0]
    print(" worker 0 MMA warp: DL arrive(epi thread)->dl_full seen %d | ->U issued %d ; el_full arrive->seen %d | ->G issued %d" % (
        np.median(x[b, 16 + 8] - x[b, 16 + 1]), np.median(x[b, 16 + 2] - x[b, 16 + 8]), np.median(x[b, 16 + 9] - x[b, 16 + 4]), np.median(x[b, 16 + 5] - x[b, 16 + 9])))
    x = t[W]
    print(" solver clock64: polled->corr %d | corr->solved %d | solved->published %d | published->next polled %d" % (
        np.median(x[b, 16 + 9] - x[b, 16 + 8]), np.median(x[b, 16 + 10] - x[b, 16 + 9]), np.median(x[b, 16 + 11] - x[b, 16 + 10]), np.median(x[b + 1, 16 + 8] - x[b, 16 + 11])))
ld = med(w[:, b, 13] - w[:, b, 12])
print(" tile gather issue->landed: min %d med %d max %d" % (ld.min(), np.median(ld), ld.max()))
slack = med(w[:, b, 5] - w[:, c, 13])
print(" tile landed before its G issue by: min %d med %d" % (slack.min(), np.median(slack)))
