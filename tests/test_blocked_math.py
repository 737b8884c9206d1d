"""The algebra the blocked sweep kernels implement, restated in numpy (float64) and checked against the plain per-marker loop.

Within a block of B markers the Gauss-Seidel steps of a linear rule (emRR Rcpp20260726ai.cpp:335, emBA :107-111, BayesRR :835 ...)
form the unit-lower-triangular system (I + A L) dE = A g + c  (bwgr_b200/csrc/common.cuh lin_coef, csrc/block_inv.cu), and a
look-ahead of D blocks takes g from a residual that is D blocks stale, corrected with the cross Gram blocks:
g_c = X_c' E_(c-D) - sum_d (X_c' X_(c-d)) dE_(c-d)  (csrc/sweep_pipe.cu).  Both are exact reformulations: this file is the
CPU proof of that (no device involved), for D = 0 .. 3 -- D = 2, 3 are the next step of DESIGN.md section 8."""
import numpy as np
import pytest

from conftest import synth


def _sequential(X, e, b, lam, kappa, order):
    e, b = e.copy(), b.copy()
    xx = (X * X).sum(0)
    for j in order:
        b1 = (X[:, j] @ e + xx[j] * b[j]) / (xx[j] + lam[j])
        e -= kappa * X[:, j] * (b1 - b[j])  # emBA applies the residual update twice (kappa = 2)
        b[j] = b1
    return b, e


def _blocked(X, e, b, lam, kappa, order, B, D, use_inverse):
    e, b = e.copy(), b.copy()
    xx = (X * X).sum(0)
    blocks = [order[s:s + B] for s in range(0, len(order), B)]
    steps = []    # dE of every finished block
    applied = 0   # blocks whose update has reached e
    for c, cols in enumerate(blocks):
        while applied < c - D:  # the workers are D blocks behind the solver
            e -= X[:, blocks[applied]] @ steps[applied]
            applied += 1
        Xc = X[:, cols]
        g = Xc.T @ e
        for d in range(applied, c):  # stale residual: correct with the cross Gram blocks
            g -= (Xc.T @ X[:, blocks[d]]) @ steps[d]
        a = kappa / (xx[cols] + lam[cols])
        cvec = -kappa * lam[cols] * b[cols] / (xx[cols] + lam[cols])
        L = np.tril(Xc.T @ Xc, -1)
        M = np.eye(len(cols)) + a[:, None] * L
        rhs = a * g + cvec
        dE = np.linalg.inv(M) @ rhs if use_inverse else np.linalg.solve(M, rhs)
        b[cols] += dE / kappa
        steps.append(dE)
    while applied < len(blocks):
        e -= X[:, blocks[applied]] @ steps[applied]
        applied += 1
    return b, e


@pytest.mark.parametrize("D", [0, 1, 2, 3])
@pytest.mark.parametrize("kappa", [1.0, 2.0])
def test_blocked_look_ahead_is_exact(D, kappa):
    X, y = synth(300, 700, seed=4)
    X = X.astype(np.float64)
    rng = np.random.default_rng(D)
    order = rng.permutation(700)
    lam = rng.uniform(20.0, 400.0, size=700)   # per-marker penalties (emBA / BayesA / emDE); a constant for emRR
    b0 = rng.normal(size=700) * 0.01
    e0 = y - y.mean() - X @ b0
    want_b, want_e = _sequential(X, e0, b0, lam, kappa, order)
    for use_inverse in (False, True):
        got_b, got_e = _blocked(X, e0, b0, lam, kappa, order, 128, D, use_inverse)
        assert np.abs(got_b - want_b).max() <= 1e-10 * np.abs(want_b).max()
        assert np.abs(got_e - want_e).max() <= 1e-9 * np.abs(want_e).max()
    if kappa == 1.0:  # with one update per marker the residual stays the residual of b
        assert np.abs((y - y.mean() - X @ want_b) - want_e).max() <= 1e-9


def test_mrr3_rotation_and_analytic_centring_are_exact():
    """MRR3's per-marker k x k solve (RcppEigen20230423.cpp:504-521, complete Y) against what the device runs:
    k independent ridge recurrences on residuals rotated by S U (S = diag(iVe)^1/2, S^-1 iG S^-1 = U Lambda U'), genotypes left
    uncentred, the column centring carried as a running shift c (e_true = e_stored + c, g = x'e_stored + c sx_j)."""
    rng = np.random.default_rng(8)
    n, p, k = 120, 60, 4
    X = rng.integers(0, 3, size=(n, p)).astype(np.float64)
    m = X.mean(0)
    Xc = X - m
    A = rng.normal(size=(k, k))
    iG = A @ A.T + k * np.eye(k)
    iVe = rng.uniform(0.5, 3.0, size=k)
    Y = rng.normal(size=(n, k))
    E0 = Y - Y.mean(0)
    b0 = rng.normal(size=(p, k)) * 0.05
    E0 = E0 - Xc @ b0
    order = rng.permutation(p)
    # the reference: one k x k system per marker on the centred column
    E, b = E0.copy(), b0.copy()
    for j in order:
        xx = Xc[:, j] @ Xc[:, j]
        rhs = (Xc[:, j] @ E + xx * b[j]) * iVe
        b1 = np.linalg.solve(iG + xx * np.diag(iVe), rhs)
        E -= np.outer(Xc[:, j], b1 - b[j])
        b[j] = b1
    # the device's form
    S = np.sqrt(iVe)
    lam, U = np.linalg.eigh(iG / np.outer(S, S))
    Et = E0 * S @ U                      # E~ = E S U   (n x k)
    bt = b0 * S @ U                      # b~_j = U' S b_j
    c = np.zeros(k)
    sx = X.sum(0)
    for j in order:
        xx = X[:, j] @ X[:, j] - sx[j] ** 2 / n          # centred xx_j = xx_j - n m_j^2
        g = X[:, j] @ Et + c * sx[j]                      # x_c'e_true with sum(e_true) = 0
        b1 = (g + xx * bt[j]) / (xx + lam)
        d = b1 - bt[j]
        Et -= np.outer(X[:, j], d)                        # workers: uncentred update
        c += m[j] * d                                     # solver: running mean shift
        bt[j] = b1
    Et += c                                               # e_true = e_stored + c
    E_back = (Et @ U.T) / S
    b_back = (bt @ U.T) / S
    assert np.abs(b_back - b).max() <= 1e-10 * np.abs(b).max()
    assert np.abs(E_back - E).max() <= 1e-10 * np.abs(E).max()


def test_spike_slab_closed_forms():
    """|e2|^2 - |e1|^2 without the two n-length temporaries (common.cuh marker_rule): BayesB / emBB compare the
    new effect with zero, KMUP / BayesDpi compare two draws; and the folded acceptance tests of the fast chain
    (u < 1/(1 + R e^x)  <=>  x < log((1/u - 1)/R);   u < min(1, q e^-x)  <=>  x < log(q/u))."""
    rng = np.random.default_rng(3)
    n = 50
    x = rng.integers(0, 3, size=n).astype(np.float64)
    e = rng.normal(size=n)
    g, xx = x @ e, x @ x
    b0, b1, b2 = 0.3, -0.2, 0.7
    n1 = ((e - x * (b1 - b0)) ** 2).sum()
    assert np.isclose(((e + x * b0) ** 2).sum() - n1, b1 * (2 * g + xx * (2 * b0 - b1)))
    assert np.isclose(((e - x * (b2 - b0)) ** 2).sum() - n1, (b2 - b1) * (-2 * g + xx * (b1 + b2 - 2 * b0)))
    for u in (0.01, 0.3, 0.77, 0.999):
        for xv in (-30.0, -1.0, 0.0, 0.4, 25.0):
            for R in (0.05, 1.0, 19.0):
                assert (u < 1 / (1 + R * np.exp(xv))) == (xv < np.log((1 / u - 1) / R))
            for q in (0.1, 0.5, 0.9):
                assert (u < min(1.0, q * np.exp(-xv))) == (xv < np.log(q / u))


def _split_limbs(q):
    """numpy image of split_limbs() in csrc/sweep_pipe.cu: four balanced signed digits, q = l0 + 2^8 l1 + 2^16 l2 + 2^24 l3."""
    q = q.astype(np.int64)
    out = []
    for _ in range(3):
        l = ((q & 0xFF) ^ 0x80) - 0x80  # (signed char)(q & 0xFF)
        out.append(l)
        q = (q - l) >> 8
    out.append(q)
    return out


def test_int8_limbs_are_exact():
    """The tensor-core passes see the residuals as four int8 limbs of a 31-bit fixed-point image: the split is
    exact and fits int8 for |q| <= 2^30, X'q is recovered exactly from the four int32 limb products, and the per-slab partial sums
    stay far inside int32 (352 rows x |x| <= 127 x |limb| <= 128)."""
    rng = np.random.default_rng(6)
    q = np.concatenate([rng.integers(-2 ** 30, 2 ** 30 + 1, size=5000), [2 ** 30, -2 ** 30, 0, 127, 128, -128, -129, 2 ** 24 - 1]])
    l0, l1, l2, l3 = _split_limbs(q)
    for l in (l0, l1, l2, l3):
        assert l.min() >= -128 and l.max() <= 127
    assert np.array_equal(l0 + (l1 << 8) + (l2 << 16) + (l3 << 24), q)
    rows = 352
    X = rng.integers(-128, 128, size=(rows, 16))
    qe = rng.integers(-2 ** 30, 2 ** 30 + 1, size=rows)
    limbs = _split_limbs(qe)
    parts = [X.T @ l for l in limbs]
    assert max(np.abs(pt).max() for pt in parts) < 2 ** 31
    assert np.array_equal(parts[0] + (parts[1] << 8) + (parts[2] << 16) + (parts[3] << 24), X.T @ qe)


def test_swizzle_128b_offsets_are_a_bijection():
    """sw128_off() of csrc/sweep_pipe.cu (K-major SWIZZLE_128B atom stack): every (row, K byte) of a 128 x 128 B tile maps to a
    distinct byte, 16-byte chunks stay whole, chunk c of row r lands in chunk c ^ (r & 7) of its 128-byte row."""
    n, kb = np.meshgrid(np.arange(128), np.arange(128), indexing="ij")
    off = (n >> 3) * 1024 + (n & 7) * 128 + ((((kb >> 4) ^ (n & 7)) & 7) << 4) + (kb & 15)
    assert sorted(off.ravel().tolist()) == list(range(128 * 128))
    assert np.array_equal(off // 128, n)                              # a row stays inside its own 128-byte line
    assert np.array_equal((off % 128) // 16, (kb >> 4) ^ (n & 7))     # XOR swizzle of the 16-byte chunk index
    assert np.array_equal(off % 16, kb % 16)


def test_self_validating_words():
    """Grid reduction without a barrier: a 64-bit word carries (value << 12) | tag, tag_of(use) = use % 4095 + 1 is
    never 0 (a zeroed buffer can never look valid) and differs between consecutive uses of a slot; 52 bits hold any partial
    sum of the sweep (|sum_l 2^(8l) X'limb| <= 2^50 by the Cauchy-Schwarz bound the host checks)."""
    tags = [(use % 4095) + 1 for use in range(3 * 4095)]
    assert min(tags) == 1 and max(tags) == 4095 and all(a != b for a, b in zip(tags, tags[1:]))
    for v in (0, 1, -1, 2 ** 50, -(2 ** 50), 123456789012345):
        for tag in (1, 77, 4095):
            w = ((v << 12) | tag) & (2 ** 64 - 1)
            assert w & 0xFFF == tag
            back = w >> 12
            if back >= 2 ** 51:
                back -= 2 ** 52
            assert back == v


def test_genotype_codes_read_as_e4m3_subnormals():
    """The exact fp8 Gram path (csrc/gram_tc.cu, kIdescF8): the int8 bytes 0..7 ARE the E4M3 subnormals code * 2^-9, every
    product code_i code_j 2^-18 is exact in fp32, and a column's sum of squares stays an exact fp32 integer while it is below
    2^24 (n = 50,000 rows of codes {0,1,2}: at most 200,000)."""
    torch = pytest.importorskip("torch")
    codes = torch.arange(0, 8, dtype=torch.uint8)
    assert torch.equal(codes.view(torch.float8_e4m3fn).float() * 512.0, codes.float())
    rng = np.random.default_rng(9)
    X = rng.integers(0, 3, size=(50000, 6)).astype(np.uint8)
    as_f8 = torch.from_numpy(X).view(torch.float8_e4m3fn).float().numpy()     # what the tensor core multiplies
    G = (as_f8.T.astype(np.float32) @ as_f8.astype(np.float32)) * np.float32(2.0 ** 18)  # fp32 accumulate, epilogue rescale
    assert np.array_equal(G.astype(np.int64), X.astype(np.int64).T @ X.astype(np.int64))


def test_philox_known_answers(tmp_path):
    """philox4x32_10() of csrc/common.cuh (the counter-based generator behind every Gibbs draw) against the known-answer vectors
    published with Random123 (Salmon et al. 2011, kat_vectors: zero, all-ones, digits-of-pi counter / key).  Host-only build of
    the same header with nvcc; no device needed."""
    import os
    import shutil
    import subprocess
    from conftest import ROOT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "kat_philox")
    subprocess.check_call([nvcc, "-std=c++17", "-w", "-I", os.path.join(ROOT, "bwgr_b200", "csrc"), "-o", exe,
                           os.path.join(ROOT, "tests", "kat_philox.cu")], stderr=subprocess.DEVNULL)
    out = subprocess.check_output([exe]).decode().split("\n")
    assert out[0] == "6627e8d5 e169c58d bc57ac4c 9b00dbd8"
    assert out[1] == "408f276d 41c83b0e a20bc7c6 6d5451fd"
    assert out[2] == "d16cfe09 94fdcceb 5001e420 24126ea1"


def test_fp4_shadow_layout_expands_to_e2m1_nibbles():
    """gram_fp4.cu: the packed 2-bit shadow (word k = rows 16k..16k+15, nibble i = {row 16k+i, row 16k+8+i}) expands with
    (w & 0x33333333) << 1 and (w >> 1) & 0x66666666 into E2M1 nibbles code << 1 (0 -> 0.0, 1 -> 1.0, 2 -> 2.0), element 2i in
    the low nibble of byte i -- and E2M1 decodes those nibbles back to the codes."""
    rng = np.random.default_rng(3)
    codes = rng.integers(0, 3, size=(64, 16)).astype(np.uint32)   # 64 words x 16 rows
    w = np.zeros(64, dtype=np.uint32)
    for r in range(16):
        w |= codes[:, r] << np.uint32(4 * (r & 7) + 2 * (r >> 3))
    lo = (w & np.uint32(0x33333333)) << np.uint32(1)
    hi = (w >> np.uint32(1)) & np.uint32(0x66666666)
    e2m1 = {0: 0.0, 1: 0.5, 2: 1.0, 3: 1.5, 4: 2.0, 5: 3.0, 6: 4.0, 7: 6.0}
    for i in range(8):
        nib_lo = (lo >> np.uint32(4 * i)) & np.uint32(15)
        nib_hi = (hi >> np.uint32(4 * i)) & np.uint32(15)
        assert np.array_equal(nib_lo, codes[:, i] << 1) and np.array_equal(nib_hi, codes[:, 8 + i] << 1)
        assert all(e2m1[int(v)] == float(c) for v, c in zip(nib_lo, codes[:, i]))
    # a code 3 is caught by the pack kernel's test (both bits of a field set)
    assert (np.uint32(3 << 6) & (np.uint32(3 << 6) >> np.uint32(1)) & np.uint32(0x55555555)) != 0
