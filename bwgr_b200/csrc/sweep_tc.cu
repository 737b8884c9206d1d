// sweep_tc.cu -- the blocked exact Gauss-Seidel / Gibbs sweep (K1 + K5 + K2 of SURVEY 2c), v2:
// both streaming contractions on the 5th-gen tensor cores, exact in fixed point.
//
// One persistent cooperative kernel per sweep.  CTA c owns the row slab [c*R, (c+1)*R) of every genotype
// column and the matching slab of the residuals E (float master copy in shared memory).  For each block
// of 128 markers in this sweep's order (Rcpp20260726ai.cpp:331):
//   1. the slab of X_B is staged ONCE in shared memory (cp.async gather, double buffered, one block ahead)
//      in the canonical SWIZZLE_128B layout [128-row atom][marker][128 B];
//   2. g_B = X_B' E:  E is quantised to 31-bit fixed point (e = q * e_q, e_q a power of two) and split in
//      four signed int8 limbs, so  g = sum_l 2^(8l) * (X_B' limb_l)  is FOUR columns of one
//      tcgen05.mma kind::i8 (A = X tile, K-major; B = limb matrix, K-major; exact int32 accumulation in
//      TMEM).  The per-CTA partial is an integer: it is added to the block accumulator in L2 with 64-bit
//      integer atomics, so the grid-wide sum is exact and independent of the grid decomposition;
//      each CTA adds (partial << 8) + 1, so every 64-bit word carries its own arrival count in the low byte
//      and readers simply poll the words they need -- no separate grid barrier, no fence;
//   3. every CTA reads the same g_B and runs the in-block solve on the Gram block
//      X_B'X_B (gram_tc.cu) held in shared memory.  Linear rules (emRR, emBA, BayesRR, BayesA, rotated
//      MRR3: de_i = a_i*(g_i - sum_{k<i} G_ik de_k) + c_i) are a unit-lower-triangular system
//      (I + A L) de = A g + c, solved in 32-marker blocks with 32x32 inverses computed BEFORE the barrier
//      (they do not depend on E); the other rules walk the scalar chain.  Either way the result is the
//      reference's Gauss-Seidel order up to float reassociation;
//   4. E_slab -= X_B_slab * dE_B: dE is quantised to int8 limbs the same way and the SAME staged tile is
//      read as the MN-major A operand (rows = M) of a second tcgen05.mma; X is read from HBM once per sweep.
#include "kernels.h"

namespace bwgr {

namespace {

constexpr int kThreads = 256;
constexpr int kAtomBytes = 128 * 128;  // one 128-row atom of the X tile: 128 markers x 128 B
constexpr uint32_t kSpin = 1u << 22;
constexpr int kGS = 132;  // row stride (floats) of the Gram block in shared memory: conflict-free LDS.128 down a column

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0; spin < kSpin; spin++) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return true;
  }
  return false;
}
// K-major SWIZZLE_128B operand: rows of 128 B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// MN-major SWIZZLE_128B operand: 128 contiguous bytes along M per K index, 8-K groups 1024 B apart
// (stride byte offset); one 128-byte M block per instruction, so the leading byte offset is unused.
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(kAtomBytes >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t idesc_i8(int N, int a_mn) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, int& v0, int& v1, int& v2, int& v3) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ long long combine_limbs(int s0, int s1, int s2, int s3) {
  return (long long)s0 + ((long long)s1 << 8) + ((long long)s2 << 16) + ((long long)s3 << 24);
}
// q (|q| <= 2^30) -> four balanced signed int8 limbs, q = l0 + 2^8 l1 + 2^16 l2 + 2^24 l3
__device__ __forceinline__ void split_limbs(int q, int& l0, int& l1, int& l2, int& l3) {
  l0 = (int)(signed char)(q & 0xFF); q = (q - l0) >> 8;
  l1 = (int)(signed char)(q & 0xFF); q = (q - l1) >> 8;
  l2 = (int)(signed char)(q & 0xFF); q = (q - l2) >> 8;
  l3 = q;
}
// byte offset of element (row n, K byte kb) inside one K-major SWIZZLE_128B atom stack
__device__ __forceinline__ uint32_t sw128_off(int n, int kb) {
  return (uint32_t)((n >> 3) * 1024 + (n & 7) * 128 + ((((kb >> 4) ^ (n & 7)) & 7) << 4) + (kb & 15));
}

struct MarkerIn { float b0, xx, vbj, a, c, ikappa; int j; float pad; };

struct TcSmem {
  uint64_t mbar_g, mbar_u;
  uint32_t tmem_base;
  int fail;
};

struct Layout {
  int R, NA, N, ns;
  size_t xs, el, dl, gs, es, mt, mk, drw, dlt, prm, rb, total;
};
__host__ __device__ inline Layout make_layout(int R, int ns, bool gibbs) {
  Layout L;
  L.R = R; L.ns = ns;
  L.NA = (R + 127) / 128;
  L.N = ((4 * ns + 15) / 16) * 16;
  size_t o = 0;
  L.xs = o; o += (size_t)2 * L.NA * kAtomBytes;            // X tiles (1024-aligned)
  L.el = o; o += (size_t)L.NA * (L.N / 8) * 1024;          // E limbs  [atom][N/8][8][128]
  L.dl = o; o += (size_t)(L.N / 8) * 1024;                 // dE limbs [N/8][8][128]
  L.gs = o; o += (size_t)128 * kGS * 4;                    // Gram block, float, padded rows
  L.es = o; o += (size_t)ns * L.NA * 128 * 4;              // E master, float [ns][NA*128]
  L.mt = o; o += (size_t)ns * 4 * 8 * kGS * 4;             // 32x32 inverses, [k/4][row][4] with padded k/4 stride
  L.mk = o; o += (size_t)2 * ns * 128 * sizeof(MarkerIn);  // per-marker inputs, double buffered
  L.drw = o; o += gibbs ? (size_t)2 * ns * 128 * sizeof(MarkerDraws) : 0;
  L.dlt = o; o += (size_t)ns * 16;                         // dE scale per system
  L.prm = o; o += (size_t)3 * 128 * 4;                     // marker ids of three consecutive blocks
  L.rb = o; o += (size_t)ns * 32 * 4;                      // broadcast buffer of the triangular solve
  L.total = o + 1024;                                      // slack for the 1024 B alignment
  return L;
}

template <int MODEL>
__global__ void __launch_bounds__(kThreads, 1) sweep_tc_kernel(SweepArgs a) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ TcSmem S;
  __shared__ SysScalars sc[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ns = a.nsys, p = a.g.p, G = gridDim.x;
  const Layout L = make_layout(a.rows_per_cta, ns, model_is_gibbs(MODEL));
  const int R = L.R, NA = L.NA, N = L.N, RS = NA * 128;
  const int row0 = blockIdx.x * R;
  const int nchunk = R >> 4;
  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* Xs = base + L.xs;
  unsigned char* EL = base + L.el;
  unsigned char* DL = base + L.dl;
  float* Gs = reinterpret_cast<float*>(base + L.gs);
  float* Es = reinterpret_cast<float*>(base + L.es);
  float* Mt = reinterpret_cast<float*>(base + L.mt);
  MarkerIn* mk = reinterpret_cast<MarkerIn*>(base + L.mk);
  MarkerDraws* drw = reinterpret_cast<MarkerDraws*>(base + L.drw);
  float* dlt = reinterpret_cast<float*>(base + L.dlt);
  int* prm = reinterpret_cast<int*>(base + L.prm);
  float* rb = reinterpret_cast<float*>(base + L.rb);

  // ---- one-time setup
  if (tid < ns) sc[tid] = a.sc[tid];
  if (tid == 0) {
    S.fail = 0;
    mbar_init(&S.mbar_g, 1);
    mbar_init(&S.mbar_u, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {  // zero the operand regions (pad rows / chunks stay zero for the whole kernel)
    uint4* z = reinterpret_cast<uint4*>(base);
    const int nz = (int)((L.gs) >> 4);
    for (int i = tid; i < nz; i += kThreads) z[i] = make_uint4(0, 0, 0, 0);
  }
  for (int s = 0; s < ns; s++)
    for (int i = tid; i < RS; i += kThreads) {
      const int r = row0 + i;
      Es[s * RS + i] = (i < R && r < a.g.ld) ? a.e[(size_t)s * a.g.ld + r] : 0.0f;
    }
  // marker ids of block blk live in prm[(blk % 3) * 128 ..]; -1 = past the end
  auto load_perm = [&](int blk) {
    if (tid < 128) {
      const int pos = blk * 128 + tid;
      prm[(blk % 3) * 128 + tid] = (blk < a.nblocks && pos < p) ? a.perm[pos] : -1;
    }
  };
  load_perm(0);
  load_perm(1);
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < (NA + 1) * N) tmem_cols <<= 1;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = S.tmem_base;
  const uint32_t idesc_g = idesc_i8(N, 0), idesc_u = idesc_i8(N, 1);

  // gather the slab of block blk: one warp per marker column, lanes = consecutive 16-byte chunks.
  // w0/nw: which warps take part (the prefetch runs beside the solve, on the warps that do not solve)
  auto issue_tile = [&](int blk, int w0, int nw) {
    if (blk < a.nblocks && warp >= w0) {
      const uint32_t dst = smem_u32(Xs + (size_t)(blk & 1) * NA * kAtomBytes);
      const int* pm = prm + (blk % 3) * 128;
      const bool two = lane + 32 < nchunk;  // slabs longer than 512 rows are not built (NA <= 4)
      const uint32_t off0 = (uint32_t)((lane >> 3) * kAtomBytes + ((lane & 7) << 4));
      const uint32_t off1 = (uint32_t)(((lane + 32) >> 3) * kAtomBytes + ((lane & 7) << 4));
      const bool in0 = lane < nchunk && row0 + 16 * lane < a.g.ld, in1 = two && row0 + 16 * (lane + 32) < a.g.ld;
      const int8_t* xbase = a.g.x8 + row0 + 16 * lane;
      for (int m = warp - w0; m < 128; m += nw) {
        const int j = pm[m];
        const uint32_t dm = dst + (uint32_t)(m * 128), sw = (uint32_t)((m & 7) << 4);
        const int8_t* col = xbase + (int64_t)(j < 0 ? 0 : j) * a.g.ld;
        if (lane < nchunk) {
          if (j >= 0 && in0) cp_async16(dm + (off0 ^ sw), col);
          else asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dm + (off0 ^ sw)), "r"(0) : "memory");
        }
        if (two) {
          if (j >= 0 && in1) cp_async16(dm + (off1 ^ sw), col + 512);
          else asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dm + (off1 ^ sw)), "r"(0) : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto issue_gram = [&](int blk) {
    if (blk < a.nblocks) {
      const float* src = a.gram + (size_t)blk * 128 * 128;
      // only the triangle the solve reads: linear rules use G[i][k], k <= i (32x32 tiles on and below the
      // diagonal); the scalar chain reads row jj to the right of the diagonal tile
      for (int idx = tid; idx < 128 * 128 / 4; idx += kThreads) {
        const int row = idx >> 5, c4 = idx & 31;
        const bool keep = model_is_linear(MODEL) ? (c4 >> 3) <= (row >> 5) : (c4 >> 3) >= (row >> 5);
        if (keep) cp_async16(smem_u32(Gs + row * kGS + 4 * c4), src + 4 * idx);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // per-marker inputs of block blk (b0, xx, per-marker variance, linear coefficients, draws): they do not
  // depend on the residuals, so they are fetched one block ahead
  auto load_markers = [&](int blk, int t0, int nt) {
    if (blk >= a.nblocks) return;
    MarkerIn* mkb = mk + (size_t)(blk & 1) * ns * 128;
    MarkerDraws* drb = drw + (size_t)(blk & 1) * ns * 128;
    const int* pm = prm + (blk % 3) * 128;
    for (int idx = t0; idx < ns * 128; idx += nt) {
      const int s = idx >> 7, m = idx & 127;
      const int j = pm[m];
      MarkerIn in = {0.0f, 1.0f, 1.0f, 0.0f, 0.0f, 1.0f, j, 0.0f};
      if (j >= 0 && !sc[s].done) {
        in.b0 = a.b[(size_t)s * p + j];
        in.xx = a.xx[j];
        in.vbj = (model_has_vbj(MODEL) && a.vbv) ? a.vbv[(size_t)s * p + j] : 1.0f;
        MarkerDraws dr;
        dr.z1 = dr.z2 = dr.u = 0.0f; dr.chi = 1.0f;
        if (model_is_gibbs(MODEL)) {
          dr = marker_draws(MODEL, (uint32_t)j, (uint32_t)sc[s].sweep, (uint32_t)(a.chain0 + s), sc[s].df, a.seed_lo, a.seed_hi);
          drb[idx] = dr;
        }
        if (model_is_linear(MODEL)) {
          const LinCoef lc = lin_coef<MODEL>(in.xx, in.b0, in.vbj, sc[s], dr);
          in.a = lc.a; in.c = lc.c; in.ikappa = 1.0f / lc.kappa;
        }
      }
      mkb[idx] = in;
    }
  };
  // residual row i of system s -> four int8 limbs in the B operand of the g pass
  auto store_limbs = [&](int s, int i, float e, float qinv, bool& bad) {
    const float sv = e * qinv;
    if (!(fabsf(sv) <= 1073741824.0f)) bad = true;
    int l0, l1, l2, l3;
    split_limbs(__float2int_rn(sv), l0, l1, l2, l3);
    unsigned char* atom = EL + (size_t)(i >> 7) * (N / 8) * 1024;
    const int kb = i & 127;
    atom[sw128_off(4 * s + 0, kb)] = (unsigned char)l0; atom[sw128_off(4 * s + 1, kb)] = (unsigned char)l1;
    atom[sw128_off(4 * s + 2, kb)] = (unsigned char)l2; atom[sw128_off(4 * s + 3, kb)] = (unsigned char)l3;
  };

  bool fail = false, lfail = false;
  const bool tracing = a.trace != nullptr && blockIdx.x == 0 && tid == 0;
#define BWGR_STAMP(k) do { if (tracing) a.trace[(size_t)blk * 16 + (k)] = clock64(); } while (0)
  issue_tile(0, 0, kThreads / 32);
  issue_gram(0);
  load_markers(0, tid, kThreads);
  for (int s = 0; s < ns; s++)
    for (int i = tid; i < R; i += kThreads) store_limbs(s, i, Es[s * RS + i], sc[s].e_qinv, fail);

#pragma unroll 1
  for (int blk = 0; blk < a.nblocks; blk++) {
    const unsigned char* Xt = Xs + (size_t)(blk & 1) * NA * kAtomBytes;
    const uint32_t par = (uint32_t)blk & 1u;
    const int nvalid = min(128, p - blk * 128);
    MarkerIn* mkb = mk + (size_t)(blk & 1) * ns * 128;
    MarkerDraws* drb = drw + (size_t)(blk & 1) * ns * 128;
    long long* gblk = a.gacc + (size_t)blk * kNC * ns * 128;  // this block's accumulators [kNC][ns][128] (zeroed by the host)

    asm volatile("cp.async.wait_group 0;" ::: "memory");   // this block's X tile and Gram block have landed
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // limbs (generic stores) -> tensor core
    if (__syncthreads_or(lfail ? 1 : 0)) break;  // uniform exit if any wait of the previous block timed out
    BWGR_STAMP(0);

    // ---- 1. g pass on the tensor core: D[marker][limb] = sum_rows X[marker][row] * limb[row]
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int at = 0; at < NA; at++) {
        const uint64_t ad = desc_k_sw128(smem_u32(Xt + (size_t)at * kAtomBytes));
        const uint64_t bd = desc_k_sw128(smem_u32(EL + (size_t)at * (N / 8) * 1024));
#pragma unroll
        for (int k4 = 0; k4 < 4; k4++) umma_i8(tmem_base, ad + (uint64_t)(2 * k4), bd + (uint64_t)(2 * k4), idesc_g, (at | k4) != 0);
      }
      umma_commit(&S.mbar_g);
    }
    BWGR_STAMP(1);
    if (warp < 4) {
      // g epilogue: TMEM -> integer partial -> L2 accumulator (thread = marker).  The word carries the
      // partial in bits 8.. and an arrival count in the low byte.
      if (!mbar_wait(&S.mbar_g, par)) lfail = true;
      BWGR_STAMP(2);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int s = 0; s < ns; s++) {
        int s0, s1, s2, s3;
        tmem_ld4(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(4 * s), s0, s1, s2, s3);
        const long long gq = combine_limbs(s0, s1, s2, s3);
        atomicAdd(reinterpret_cast<unsigned long long*>(gblk) + ((size_t)(blockIdx.x % kNC) * ns + s) * 128 + tid, (unsigned long long)((gq << 8) + 1));
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else if (model_is_linear(MODEL)) {
      // ---- 2a. invert the four 32x32 diagonal blocks of I + A L (E-independent) beside the g pass
      for (int task = warp - 4; task < ns * 4; task += 4) {
        const int s = task >> 2, d = task & 3;
        const MarkerIn* mks = mkb + s * 128 + 32 * d;
        float x[32];
#pragma unroll
        for (int i = 0; i < 32; i++) {
          float acc0 = 0.0f, acc1 = 0.0f;
          const float4* grow = reinterpret_cast<const float4*>(Gs + (32 * d + i) * kGS + 32 * d);
#pragma unroll
          for (int k4 = 0; k4 < (i + 3) / 4; k4++) {
            const float4 gv = grow[k4];
            if (4 * k4 + 0 < i) acc0 = fmaf(gv.x, x[4 * k4 + 0], acc0);
            if (4 * k4 + 1 < i) acc1 = fmaf(gv.y, x[4 * k4 + 1], acc1);
            if (4 * k4 + 2 < i) acc0 = fmaf(gv.z, x[4 * k4 + 2], acc0);
            if (4 * k4 + 3 < i) acc1 = fmaf(gv.w, x[4 * k4 + 3], acc1);
          }
          x[i] = (i == lane) ? 1.0f : -mks[i].a * (acc0 + acc1);
        }
        // M[i][c] (c = lane) stored as Mt4[c/4][i][c%4] with a padded c/4 stride: conflict-free both ways
        float* mt = Mt + (size_t)(s * 4 + d) * 8 * kGS + (lane >> 2) * kGS + (lane & 3);
#pragma unroll
        for (int i = 0; i < 32; i++) mt[4 * i] = x[i];
      }
    }
    BWGR_STAMP(3);
    __syncthreads();
    BWGR_STAMP(4);

    // ---- 2b. solve warps poll their accumulator words and run the in-block solve; the other warps
    //          prefetch the next block (marker ids, genotype tile, per-marker inputs) meanwhile
    const int nsolve = ns < kThreads / 32 ? ns : kThreads / 32;
    if (warp >= nsolve || ns > kThreads / 32) {
      if (warp >= nsolve) {
        const int t0 = (warp - nsolve) * 32 + lane, nt = (kThreads / 32 - nsolve) * 32;
        if (blk + 2 < a.nblocks + 2)
          for (int m = t0; m < 128; m += nt) {
            const int pos = (blk + 2) * 128 + m;
            prm[((blk + 2) % 3) * 128 + m] = (blk + 2 < a.nblocks && pos < p) ? a.perm[pos] : -1;
          }
        issue_tile(blk + 1, nsolve, kThreads / 32 - nsolve);
        load_markers(blk + 1, t0, nt);
      }
    }
    for (int s = warp; s < ns; s += kThreads / 32) {
      const SysScalars Sy = sc[s];
      const MarkerIn* mks = mkb + s * 128;
      float g[4], de[4];
      {
        // every word is (sum of partials << 8) + arrivals; copy c collects the CTAs with blockIdx % kNC == c
        long long q[4] = {0, 0, 0, 0};
        uint32_t spins = 0;
#pragma unroll 1
        for (int c = 0; c < kNC; c++) {
          const long long* gq = gblk + ((size_t)c * ns + s) * 128;
          const int expect = (G - c + kNC - 1) / kNC;
          long long w[4];
          while (true) {
            bool ok = true;
#pragma unroll
            for (int t = 0; t < 4; t++) {
              asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(w[t]) : "l"(gq + 32 * t + lane) : "memory");
              ok = ok && ((int)(w[t] & 0xFF) == expect);
            }
            if (__all_sync(0xffffffffu, ok)) break;
            if (++spins > kSpin) { lfail = true; atomicExch(a.err, 3); break; }
          }
#pragma unroll
          for (int t = 0; t < 4; t++) q[t] += w[t] >> 8;
        }
        BWGR_STAMP(5);
#pragma unroll
        for (int t = 0; t < 4; t++) {
          g[t] = (float)((double)q[t] * (double)Sy.e_q);
          de[t] = 0.0f;
        }
      }
      float nb[4] = {0.f, 0.f, 0.f, 0.f}, nd[4] = {1.f, 1.f, 1.f, 1.f}, nv[4] = {1.f, 1.f, 1.f, 1.f};
      if (Sy.done) {
        // converged system (emEN): no update
      } else if (model_is_linear(MODEL)) {
        float r[4];
#pragma unroll
        for (int t = 0; t < 4; t++) r[t] = fmaf(mks[32 * t + lane].a, g[t], mks[32 * t + lane].c);
        float* rbs = rb + s * 32;
#pragma unroll
        for (int d = 0; d < 4; d++) {
          // de_d = M_d r_d : broadcast r_d through shared memory, 128-bit loads down the k axis
          __syncwarp();
          rbs[lane] = r[d];
          __syncwarp();
          const float* mt = Mt + (size_t)(s * 4 + d) * 8 * kGS + 4 * lane;
          float ac[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int k4 = 0; k4 < 8; k4++) {
            const float4 mv = *reinterpret_cast<const float4*>(mt + k4 * kGS);
            const float4 rv = *reinterpret_cast<const float4*>(rbs + 4 * k4);
            ac[0] = fmaf(mv.x, rv.x, ac[0]); ac[1] = fmaf(mv.y, rv.y, ac[1]);
            ac[2] = fmaf(mv.z, rv.z, ac[2]); ac[3] = fmaf(mv.w, rv.w, ac[3]);
          }
          const float acc = (ac[0] + ac[1]) + (ac[2] + ac[3]);
          de[d] = acc;
          if (d < 3) {
            __syncwarp();
            rbs[lane] = acc;
            __syncwarp();
#pragma unroll
            for (int d2 = 0; d2 < 4; d2++) {
              if (d2 > d) {
                // r_d2 -= a * sum_k G[32 d2 + lane][32 d + k] * de_d[k]
                const float* grow = Gs + (32 * d2 + lane) * kGS + 32 * d;
                float fa[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k4 = 0; k4 < 8; k4++) {
                  const float4 gv = *reinterpret_cast<const float4*>(grow + 4 * k4);
                  const float4 dv = *reinterpret_cast<const float4*>(rbs + 4 * k4);
                  fa[0] = fmaf(gv.x, dv.x, fa[0]); fa[1] = fmaf(gv.y, dv.y, fa[1]);
                  fa[2] = fmaf(gv.z, dv.z, fa[2]); fa[3] = fmaf(gv.w, dv.w, fa[3]);
                }
                r[d2] = fmaf(-mks[32 * d2 + lane].a, (fa[0] + fa[1]) + (fa[2] + fa[3]), r[d2]);
              }
            }
          }
        }
      } else {
#pragma unroll
        for (int t = 0; t < 4; t++) {
#pragma unroll 8
          for (int i = 0; i < 32; i++) {
            const int jj = 32 * t + i;
            if (jj >= nvalid) break;
            const float gc = __shfl_sync(0xffffffffu, g[t], i);
            const MarkerIn in = mks[jj];
            MarkerDraws dr;
            if (model_is_gibbs(MODEL)) dr = drb[s * 128 + jj];
            else { dr.z1 = dr.z2 = dr.u = 0.0f; dr.chi = 1.0f; }
            const RuleOut ro = marker_rule<MODEL>(gc, in.xx, in.b0, in.vbj, Sy, dr);
            if (lane == i) { nb[t] = ro.b; nd[t] = ro.d; nv[t] = ro.vbj; de[t] = ro.de; }
            const float* grow = Gs + jj * kGS + lane;
#pragma unroll
            for (int tt = 0; tt < 4; tt++)
              if (tt >= t) g[tt] = fmaf(-grow[32 * tt], ro.de, g[tt]);
          }
        }
      }
      BWGR_STAMP(6);
      // quantise dE to 31-bit fixed point relative to the block maximum (all CTAs compute the same scale)
      float mx = fmaxf(fmaxf(fabsf(de[0]), fabsf(de[1])), fmaxf(fabsf(de[2]), fabsf(de[3])));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      int ex = 0;
      if (mx > 0.0f && mx < 3.0e38f) frexpf(mx, &ex);
      if (ex < -90) ex = -90;
      const float dq = ldexpf(1.0f, ex - 30), dqinv = ldexpf(1.0f, 30 - ex);
      if (!(mx < 3.0e38f)) fail = true;
      if (lane == 0) dlt[s * 4] = dq;  // scale for the update epilogue
#pragma unroll
      for (int t = 0; t < 4; t++) {
        const int jj = 32 * t + lane;
        const bool valid = jj < nvalid && !Sy.done;
        const int q = valid ? __float2int_rn(de[t] * dqinv) : 0;
        const float deq = (float)q * dq;  // the step actually applied to E, rounded to float (a 31-bit q keeps its leading 24 bits; the residual update uses the same q, so b and E stay consistent to that rounding)
        int l0, l1, l2, l3;
        split_limbs(q, l0, l1, l2, l3);
        DL[sw128_off(4 * s + 0, jj)] = (unsigned char)l0; DL[sw128_off(4 * s + 1, jj)] = (unsigned char)l1;
        DL[sw128_off(4 * s + 2, jj)] = (unsigned char)l2; DL[sw128_off(4 * s + 3, jj)] = (unsigned char)l3;
        if (valid) {
          const MarkerIn in = mks[jj];
          float bnew, dnew = nd[t], vnew = nv[t];
          if (model_is_linear(MODEL)) {
            bnew = fmaf(deq, in.ikappa, in.b0);
            if (MODEL == M_EMBA) vnew = (Sy.Sb + bnew * bnew) / (Sy.df + 1.0f);
            if (MODEL == M_BA || MODEL == M_BL) vnew = (Sy.Sb + bnew * bnew) / drb[s * 128 + jj].chi;
          } else {
            bnew = nb[t];
          }
          if (blockIdx.x == 0) {
            a.b[(size_t)s * p + in.j] = bnew;
            if (model_has_d(MODEL) && a.d) a.d[(size_t)s * p + in.j] = dnew;
            if (model_rule_writes_vbj(MODEL) && a.vbv) a.vbv[(size_t)s * p + in.j] = vnew;
          }
        }
      }
    }
    if (ns >= kThreads / 32) {  // every warp solved: nobody prefetched beside the solve, do it now
      __syncthreads();
      load_perm(blk + 2);
      issue_tile(blk + 1, 0, kThreads / 32);
      __syncthreads();
      load_markers(blk + 1, tid, kThreads);
    }
    BWGR_STAMP(7);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    BWGR_STAMP(8);

    // ---- 3. update pass on the tensor core: D[row][limb] = sum_markers X[row][marker] * dE_limb[marker]
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t bd = desc_k_sw128(smem_u32(DL));
      for (int at = 0; at < NA; at++) {
        const uint64_t ad = desc_mn_sw128(smem_u32(Xt + (size_t)at * kAtomBytes));
#pragma unroll
        for (int k4 = 0; k4 < 4; k4++)
          umma_i8(tmem_base + (uint32_t)((at + 1) * N), ad + (uint64_t)(k4 * (4096 >> 4)), bd + (uint64_t)(2 * k4), idesc_u, k4 != 0);
      }
      umma_commit(&S.mbar_u);
    }
    issue_gram(blk + 1);  // Gs is free again (solve done)
    if (!mbar_wait(&S.mbar_u, par)) lfail = true;
    BWGR_STAMP(9);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int at = warp >> 2; at < NA; at += 2) {
      const int ra = (warp & 3) * 32 + lane;  // row inside the atom = TMEM lane
      const int i = at * 128 + ra;
      for (int s = 0; s < ns; s++) {
        int s0, s1, s2, s3;
        tmem_ld4(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((at + 1) * N + 4 * s), s0, s1, s2, s3);
        const long long uq = combine_limbs(s0, s1, s2, s3);
        if (i < R) {
          const float e = Es[s * RS + i] - (float)((double)uq * (double)dlt[s * 4]);
          Es[s * RS + i] = e;
          store_limbs(s, i, e, sc[s].e_qinv, fail);  // B operand of the next block's g pass
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    BWGR_STAMP(10);
  }
#undef BWGR_STAMP
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (fail) atomicExch(a.err, 4);
  if (lfail) atomicCAS(a.err, 0, 2);
  for (int s = 0; s < ns; s++)
    for (int i = tid; i < R; i += kThreads) {
      const int r = row0 + i;
      if (r < a.g.ld) a.e[(size_t)s * a.g.ld + r] = Es[s * RS + i];
    }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

template <int MODEL>
void launch_model(const SweepArgs& a, int grid, cudaStream_t st) {
  const Layout L = make_layout(a.rows_per_cta, a.nsys, model_is_gibbs(MODEL));
  cudaFuncSetAttribute(sweep_tc_kernel<MODEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
  SweepArgs args = a;
  void* params[] = {&args};
  cudaLaunchCooperativeKernel((void*)sweep_tc_kernel<MODEL>, dim3(grid), dim3(kThreads), params, L.total, st);
}

}  // namespace

size_t sweep_blocked_smem(int rows_per_cta, int nsys) { return make_layout(rows_per_cta, nsys, true).total + 2048; }

void launch_sweep_blocked(const SweepArgs& a, int grid, cudaStream_t st) {
  switch (rule_model(a.model)) {
    case M_EMRR: launch_model<M_EMRR>(a, grid, st); break;
    case M_EMBA: launch_model<M_EMBA>(a, grid, st); break;
    case M_EMBB: launch_model<M_EMBB>(a, grid, st); break;
    case M_EMBC: launch_model<M_EMBC>(a, grid, st); break;
    case M_EMBL: launch_model<M_EMBL>(a, grid, st); break;
    case M_EMEN: launch_model<M_EMEN>(a, grid, st); break;
    case M_BRR: launch_model<M_BRR>(a, grid, st); break;
    case M_BA: launch_model<M_BA>(a, grid, st); break;
    case M_BB: launch_model<M_BB>(a, grid, st); break;
    case M_BC: launch_model<M_BC>(a, grid, st); break;
    case M_KMUP: launch_model<M_KMUP>(a, grid, st); break;
    case M_MRR: launch_model<M_MRR>(a, grid, st); break;
    case M_EMDE: launch_model<M_EMDE>(a, grid, st); break;
    case M_LASSO: launch_model<M_LASSO>(a, grid, st); break;
    case M_BL: launch_model<M_BL>(a, grid, st); break;
    case M_BDPI: launch_model<M_BDPI>(a, grid, st); break;
    default: break;
  }
}

}  // namespace bwgr
