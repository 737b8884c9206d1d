#!/usr/bin/env python
"""Summarise a BWGR_TRACE file of the clustered sweep with the FAST worker: [cta][block][32] = 16 globaltimer (ns) + 16 clock64 stamps.
CTA 8c = solver of cluster c (8 h read, 9 corrected, 10 solved, 11 published; 4 = comm warps: grid sum ready),
CTA 8c + r = workers; epilogue warp 1 lane 0: 0 step received, 1 DL written, 5 tile c landed, 3 first u_done, 6 first unit done,
4 all units done, 8 next tile's gather issued, 7 partial sent; issuer: 9 dl_full seen, 2 all atoms issued."""
import sys
import numpy as np
raw = np.fromfile(sys.argv[1], dtype=np.int64)
G, nb, K, D = raw[:4]
t = raw[4:].reshape(G, nb, K).astype(np.float64)
lo, hi = 50, min(350, nb - 4)
b = np.arange(lo, hi)
c = b + 1 + D
S = t[0]
ck = 16
print("grid", G, "blocks", nb, "D", D)
print("solver period: %.0f ns (globaltimer), %.0f cycles" % (np.median(np.diff(S[lo:hi, 11])), np.median(np.diff(S[lo:hi, ck + 11]))))
print("solver cycles: grid sum ready(comm) -> h read %d | h read -> corrected %d | corrected -> solved %d | solved -> published %d | published -> next grid sum ready %d" % (
    np.median(S[b, ck + 8] - S[b, ck + 4]), np.median(S[b, ck + 9] - S[b, ck + 8]), np.median(S[b, ck + 10] - S[b, ck + 9]),
    np.median(S[b, ck + 11] - S[b, ck + 10]), np.median(S[b + 1, ck + 4] - S[b, ck + 11])))
for wk in (1, 7, 8 * (G // 16) + 3, G - 1):
    x = t[wk]
    f = lambda k1, b1, k0, b0: np.median(x[b1, ck + k1] - x[b0, ck + k0])
    print("worker cta %3d cycles: recv->DL %d | DL->issuer sees dl_full %d | ->all atoms issued %d || recv->tile c landed %d | ->first u_done %d | ->first unit done %d | ->all units %d | ->gather issued %d | ->partial sent %d | sent->next recv %d  (busy %d)" % (
        wk, f(1, b, 0, b), f(9, b, 1, b), f(2, b, 9, b), f(5, b, 0, b), f(3, b, 5, b), f(6, b, 3, b), f(4, b, 6, b), f(8, b, 4, b), f(7, c, 8, b), f(0, b + 1, 7, c), f(7, c, 0, b)))
w = t[1:8]
print("cluster 0 (ns, globaltimer): published -> step received by its workers: med %.0f max %.0f" % (np.median(w[:, b, 0] - S[b, 11][None, :]), np.median((w[:, b, 0] - S[b, 11][None, :]).max(0))))
allw = np.array([t[i] for i in range(G) if i % 8 != 0])
last_sent = allw[:, c, 7].max(0)
print("last partial sent by ANY worker -> grid sum ready at solver 0: med %.0f ns ; last partial of cluster 0 -> ready: %.0f ns" % (
    np.median(S[c, 4] - last_sent), np.median(S[c, 4] - w[:, c, 7].max(0))))
sent = allw[:, c, 7]
print("partial sent (block c=b+1+D) relative to step received (block b): med %.0f ns, slowest worker med %.0f ns" % (np.median(sent - allw[:, b, 0]), np.median((sent - allw[:, b, 0]).max(0))))
