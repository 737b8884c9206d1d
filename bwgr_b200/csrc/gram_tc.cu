// gram_tc.cu -- Gram blocks X_B' X_B on the 5th-gen tensor cores (K6 of SURVEY 2c; not in the reference,
// required by the blocked exact Gauss-Seidel reformulation).
//
// For every block of 128 markers (in this sweep's order) G = X_B' X_B is a 128 x 128 x n int8 GEMM with
// exact int32 accumulation: tcgen05.mma kind::i8, M = N = 128, K = 32 per instruction, accumulator in
// TMEM.  A and B are the SAME shared-memory tile (marker-major, K = rows contiguous), staged as the
// canonical K-major SWIZZLE_128B layout: row m of the tile holds 128 consecutive genotype rows of marker m,
// 16-byte chunk c stored at chunk (c ^ (m & 7)); 8-row groups are 1024 B apart (SBO = 1024).
// Columns of a block are scattered in HBM when the order is shuffled (emRR & co, Rcpp20260726ai.cpp:331),
// so the tile is gathered with 16-byte cp.async (8 consecutive lanes = one 128-byte line of one column).
//
// Warp roles (416 threads): warps 0-3 gather (cp.async producer; with the packed source warps 9-12 join them), warp 4
// issues the MMAs (one elected lane) and owns the TMEM allocation, warps 5-8 drain the accumulator (tcgen05.ld -> st.global).
// Two accumulator stages (2 x 128 TMEM columns) overlap the drain of block i with the MMAs of block i+1.
// One CTA per SM, persistent over blocks.  HBM traffic: every genotype byte is read exactly once.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "kernels.h"

namespace bwgr {

namespace {

constexpr int kTileBytes = 128 * 128;
// NBAND = 1: diagonal blocks X_b'X_b only.  NBAND = 2: [X_b'X_b | X_{b-1}'X_b] (one-block look-ahead of the
// pipelined sweep, sweep_pipe.cu): row r of block b holds NBAND*128 entries, x_{b,r}'X_b then x_{b-1,r}'X_b.
// The gather is latency-bound: keep (almost) the whole 227 KB of shared memory in flight.
// A stage holds kSub consecutive 128-row atoms of every column of the band, so one visit to a column reads kSub*128
// contiguous bytes (DRAM pages are reused instead of re-opened per 128 B).
template <int NBAND> struct GramCfg {
  static constexpr int kSub = NBAND == 1 ? 4 : 2;
  static constexpr int kStages = 3;
  static constexpr int kLag = 2;
};
constexpr uint32_t kSpinLimit = 1u << 22;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one elected lane of a converged warp: the form ptxas turns into back-to-back UTC*MMA issue (a plain `lane == 0`
// branch makes it wrap every tcgen05.mma in a per-lane loop, ~250 cycles per instruction)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a protocol bug must not hang the GPU (it sets *err and lets the kernel drain).
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* err) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0; spin < kSpinLimit; spin++) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return true;
  }
  atomicExch(err, 2);
  return false;
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA tile::gather4: four rows (= four marker columns, by index) of the 2-D genotype tensor [p][ld], 128 bytes each from
// row offset crd0, land as four consecutive 128-byte rows at dst in the SWIZZLE_128B pattern the UMMA descriptor expects.
// An index outside [0, p) is filled with zeros (and still counts its bytes on the mbarrier).
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* tmap, int crd0, int i0, int i1, int i2, int i3,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(crd0), "r"(i0), "r"(i1), "r"(i2), "r"(i3), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);  // start address, 16 B units
  d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major) = 16 B
  d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}

// kind::i8, D = S32, A = B = signed int8, both K-major, M = 128, N = 128.
constexpr uint32_t kIdescI8 = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

// kind::f8f6f4 with both operands E4M3 and an F32 accumulator.  Genotype codes 0..7 stored as int8 are, read as E4M3,
// the subnormals code * 2^-9: the SAME bytes feed this (4x faster) instruction, every product code_i*code_j*2^-18 is
// exact, and so is the fp32 sum while max_j xx_j < 2^24 -- the epilogue rescales by 2^18 and lands on the exact integer.
constexpr uint32_t kIdescF8 = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int NBAND>
struct GramSmem {
  uint64_t full[GramCfg<NBAND>::kStages], empty[GramCfg<NBAND>::kStages], tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
  float sxc[128];  // column sums of this block's markers (centred Gram)
};

template <int NBAND, bool FP8, int PROD>
__global__ void __launch_bounds__(416, 1) gram_tc_kernel(GenoView g, const int* __restrict__ perm, int nblocks,
                                                         int32_t* __restrict__ gram, int out_f32, int* err,
                                                         const float* __restrict__ sx, float inv_n, int dbg,
                                                         const __grid_constant__ CUtensorMap tmap) {
  constexpr int kStages = GramCfg<NBAND>::kStages;
  constexpr int kLag = GramCfg<NBAND>::kLag;  // cp.async groups kept in flight per producer thread
  constexpr int kSub = GramCfg<NBAND>::kSub;
  constexpr int kStageBytes = NBAND * kSub * kTileBytes;
  constexpr uint32_t kAccCols = NBAND * 128;  // TMEM columns of one accumulator stage
  using Sm = GramSmem<NBAND>;
  extern __shared__ unsigned char smem_raw[];
  // tiles first (1024 B aligned for SWIZZLE_128B), bookkeeping after them
  unsigned char* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Sm* S = reinterpret_cast<Sm*>(tiles + kStages * kStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkt = (int)(g.ld >> 7);  // K tiles of 128 rows (ld is a multiple of 128)
  const int nkg = (nkt + GramCfg<NBAND>::kSub - 1) / GramCfg<NBAND>::kSub;  // stages per block

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; s++) { mbar_init(&S->full[s], PROD == 1 ? 1 : PROD == 2 ? 256 : 128); mbar_init(&S->empty[s], 1); }
    for (int s = 0; s < 2; s++) { mbar_init(&S->tmem_full[s], 1); mbar_init(&S->tmem_empty[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S->tmem_base)), "r"(2u * kAccCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = S->tmem_base;

  if (PROD == 1 && warp < 4) {
    // ===================== producer: TMA gather4 (one warp; lane l brings markers 4l..4l+3 of every tile) =====================
    if (warp == 0) {
      uint32_t it = 0;
      bool ok = true;
      for (int blk = blockIdx.x; blk < nblocks && ok; blk += gridDim.x) {
        int ids[NBAND][4];
#pragma unroll
        for (int d = 0; d < NBAND; d++)
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const int pos = (blk - d) * 128 + 4 * lane + q;
            ids[d][q] = (pos >= 0 && pos < g.p) ? perm[pos] : g.p;  // g.p = out of bounds -> zeros
          }
        for (int kg = 0; kg < nkg && ok; kg++, it++) {
          const uint32_t stage = it % kStages, phase = (it / kStages) & 1u;
          ok = mbar_wait(&S->empty[stage], phase ^ 1u, err);
          const uint32_t tbase = smem_u32(tiles + stage * kStageBytes);
          const int nsub = min(kSub, nkt - kg * kSub);
          if (lane == 0) mbar_arrive_expect_tx(&S->full[stage], (uint32_t)(nsub * NBAND * kTileBytes));
          __syncwarp();
#pragma unroll
          for (int sub = 0; sub < kSub; sub++) {
            if (sub < nsub) {
              const int kt = kg * kSub + sub;
#pragma unroll
              for (int d = 0; d < NBAND; d++)
                tma_gather4(tbase + (sub * NBAND + d) * kTileBytes + lane * 512, &tmap, kt * 128, ids[d][0], ids[d][1], ids[d][2], ids[d][3],
                            &S->full[stage]);
            }
          }
        }
      }
    }
  } else if (PROD == 2 && (warp < 4 || warp >= 9)) {
    // ===================== producer: 2-bit packed source (codes 0..3), expanded to bytes in flight =====================
    // The gather is what bounds this kernel (~4.3 TB/s of cp.async at 50k x 50k whatever the MMA kind); reading the packed
    // shadow copy moves 4x fewer bytes.  A 16-byte packed chunk = 64 rows of one marker = half a 128-row atom: it is
    // loaded into registers one stage ahead, expanded with two shift/mask steps per word and stored as four swizzled
    // 16-byte chunks.  Stage = kSub atoms, i.e. CPC = 2*kSub packed chunks per marker.
    // Eight producer warps (0-3 and 9-12): the expansion, not the gather, is what this producer costs.
    constexpr int CPC = 2 * kSub;          // packed 16-byte chunks per marker per stage
    constexpr int NIT = CPC / 2;           // chunks per thread per band per stage (128 markers * CPC chunks / 256 threads)
    constexpr int NLD = NIT * NBAND;
    constexpr int MSTEP = 256 / CPC;       // marker stride between a thread's chunks
    const int t = warp < 4 ? threadIdx.x : threadIdx.x - 160;  // 0..255
    const int c = t % CPC, m0 = t / CPC;   // this thread's chunk within the stage and its first marker; markers m0 + MSTEP*i
    uint32_t it = 0;
    bool ok = true;
    uint4 cur[NLD], nxt[NLD];
    auto load_stage = [&](const uint8_t* const (&colp)[NBAND][NIT], const bool (&val)[NBAND][NIT], int kg, uint4 (&dst)[NLD]) {
#pragma unroll
      for (int d = 0; d < NBAND; d++)
#pragma unroll
        for (int i = 0; i < NIT; i++) {
          const int64_t off = ((int64_t)kg * CPC + c) * 16;  // byte offset inside the packed column
          dst[d * NIT + i] = (val[d][i] && off < g.ldb) ? __ldg(reinterpret_cast<const uint4*>(colp[d][i] + off)) : make_uint4(0, 0, 0, 0);
        }
    };
    for (int blk = blockIdx.x; blk < nblocks && ok; blk += gridDim.x) {
      const uint8_t* colp[NBAND][NIT];
      bool val[NBAND][NIT];
#pragma unroll
      for (int d = 0; d < NBAND; d++)
#pragma unroll
        for (int i = 0; i < NIT; i++) {
          const int m = m0 + MSTEP * i;
          const int pos = (blk - d) * 128 + m;
          val[d][i] = pos >= 0 && pos < g.p;
          colp[d][i] = g.x2 + (int64_t)(val[d][i] ? perm[pos] : 0) * g.ldb;
        }
      load_stage(colp, val, 0, cur);
      for (int kg = 0; kg < nkg && ok; kg++, it++) {
        if (kg + 1 < nkg) load_stage(colp, val, kg + 1, nxt);
        const uint32_t stage = it % kStages, phase = (it / kStages) & 1u;
        ok = mbar_wait(&S->empty[stage], phase ^ 1u, err);
        const uint32_t tbase = smem_u32(tiles + stage * kStageBytes);
        const int sub = c >> 1;  // atom of this chunk; rows 64*(c&1) .. +63 inside it -> output chunks 4*(c&1) .. +3
#pragma unroll
        for (int d = 0; d < NBAND; d++)
#pragma unroll
          for (int i = 0; i < NIT; i++) {
            const int m = m0 + MSTEP * i;
            const uint4 pk = cur[d * NIT + i];
            const uint32_t w4[4] = {pk.x, pk.y, pk.z, pk.w};
            const uint32_t rowb = tbase + (sub * NBAND + d) * kTileBytes + m * 128;
#pragma unroll
            for (int q = 0; q < 4; q++) {  // packed word q = rows 16q .. 16q+15 of the half atom = one 16-byte output chunk
              // interleaved shadow layout (geno.cu): bytes of rows 4r .. 4r+3 = (w >> 2r) & 0x03030303
              const uint32_t oc = (uint32_t)(4 * (c & 1) + q);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowb + ((oc ^ ((uint32_t)m & 7u)) << 4)),
                           "r"(w4[q] & 0x03030303u), "r"((w4[q] >> 2) & 0x03030303u), "r"((w4[q] >> 4) & 0x03030303u),
                           "r"((w4[q] >> 6) & 0x03030303u) : "memory");
            }
          }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&S->full[stage]);
        if (kg + 1 < nkg) {
#pragma unroll
          for (int q = 0; q < NLD; q++) cur[q] = nxt[q];
        }
      }
    }
  } else if (warp < 4) {
    // ===================== producer: gather tiles with cp.async (fallback when no tensor map could be built) =====================
    const int t = threadIdx.x;  // 0..127
    uint32_t it = 0;            // tile counter (runs over blocks and K tiles)
    bool ok = true;
    for (int blk = blockIdx.x; blk < nblocks && ok; blk += gridDim.x) {
      const int8_t* colp[NBAND][8];
      uint32_t nbytes[NBAND][8];
#pragma unroll
      for (int d = 0; d < NBAND; d++)
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int m = i * 16 + (t >> 3);
          const int pos = (blk - d) * 128 + m;
          const bool valid = pos >= 0 && pos < g.p;
          const int j = valid ? perm[pos] : 0;
          colp[d][i] = g.x8 + (int64_t)j * g.ld + ((t & 7) << 4);
          nbytes[d][i] = valid ? 16u : 0u;
        }
      for (int kg = 0; kg < nkg && ok; kg++, it++) {
        const uint32_t stage = it % kStages, phase = (it / kStages) & 1u;
        ok = mbar_wait(&S->empty[stage], phase ^ 1u, err);
        const uint32_t tbase = smem_u32(tiles + stage * kStageBytes);
#pragma unroll
        for (int sub = 0; sub < kSub; sub++) {
          const int kt = kg * kSub + sub;
          if (kt < nkt) {
#pragma unroll
            for (int d = 0; d < NBAND; d++)
#pragma unroll
              for (int i = 0; i < 8; i++) {
                const int m = i * 16 + (t >> 3);
                const uint32_t dst = tbase + (sub * NBAND + d) * kTileBytes + m * 128 + ((((uint32_t)t & 7u) ^ ((uint32_t)m & 7u)) << 4);
                cp_async16_zfill(dst, colp[d][i] + (int64_t)kt * 128, nbytes[d][i]);
              }
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (it >= kLag) {
          asm volatile("cp.async.wait_group %0;" ::"n"(kLag) : "memory");
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive(&S->full[(it - kLag) % kStages]);
        }
      }
    }
    // drain: signal the last (up to kLag) tiles
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (uint32_t q = (it > kLag ? it - kLag : 0); q < it; q++) mbar_arrive(&S->full[q % kStages]);
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    uint32_t it = 0, bi = 0;
    bool ok = true;
    for (int blk = blockIdx.x; blk < nblocks && ok; blk += gridDim.x, bi++) {
      const uint32_t as = bi & 1u, aphase = (bi >> 1) & 1u;
      ok = mbar_wait(&S->tmem_empty[as], aphase ^ 1u, err);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tmem_d = tmem_base + as * kAccCols;
      for (int kg = 0; kg < nkg && ok; kg++, it++) {
        const uint32_t stage = it % kStages, phase = (it / kStages) & 1u;
        ok = mbar_wait(&S->full[stage], phase, err);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
#pragma unroll
          for (int sub = 0; sub < kSub; sub++) {
            const int kt = kg * kSub + sub;
            if (kt < nkt && !(dbg & 1)) {  // dbg bit 0 (BWGR_GRAM_DBG=1): gather only, no MMAs -- isolates the gather rate
              const uint64_t desc = make_desc_sw128(smem_u32(tiles + stage * kStageBytes + sub * NBAND * kTileBytes));
#pragma unroll
              for (int d = 0; d < NBAND; d++) {
                const uint64_t bdesc = make_desc_sw128(smem_u32(tiles + stage * kStageBytes + (sub * NBAND + d) * kTileBytes));
#pragma unroll
                for (int k4 = 0; k4 < 4; k4++) {
                  // band d >= 1 is stored transposed: row r = marker r of block b-d, columns = markers of block b, so that
                  // the sweep's correction threads (one per marker of block b) read it coalesced
                  const uint32_t acc = (kt | k4) != 0 ? 1u : 0u;
                  if (FP8) {
                    if (d == 0) umma_f8(tmem_d, desc + (uint64_t)(k4 * 2), desc + (uint64_t)(k4 * 2), kIdescF8, acc);
                    else umma_f8(tmem_d + (uint32_t)d * 128u, bdesc + (uint64_t)(k4 * 2), desc + (uint64_t)(k4 * 2), kIdescF8, acc);
                  } else {
                    if (d == 0) umma_i8(tmem_d, desc + (uint64_t)(k4 * 2), desc + (uint64_t)(k4 * 2), kIdescI8, acc);
                    else umma_i8(tmem_d + (uint32_t)d * 128u, bdesc + (uint64_t)(k4 * 2), desc + (uint64_t)(k4 * 2), kIdescI8, acc);
                  }
                }
              }
            }
          }
          umma_commit(&S->empty[stage]);  // frees the stage when these MMAs have read it
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&S->tmem_full[as]);
      __syncwarp();
    }
  } else if (warp <= 8) {
    // ===================== epilogue: TMEM -> registers -> HBM =====================
    const int quarter = warp & 3;  // TMEM lanes this warp may touch
    const int row = quarter * 32 + lane;
    uint32_t bi = 0;
    bool ok = true;
    for (int blk = blockIdx.x; blk < nblocks && ok; blk += gridDim.x, bi++) {
      const uint32_t as = bi & 1u, aphase = (bi >> 1) & 1u;
      ok = mbar_wait(&S->tmem_full[as], aphase, err);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      int32_t* out = gram + ((size_t)blk * 128 + row) * (NBAND * 128);
      // centred Gram (MRR3 centres every column, RcppEigen20230423.cpp:378-379): x_ci'x_ck = x_i'x_k - sx_i sx_k / n
      float sxr0 = 0.0f, sxr1 = 0.0f;
      if (sx) {
        const int pos0 = blk * 128 + row, pos1 = (blk - 1) * 128 + row;
        sxr0 = pos0 < g.p ? sx[perm[pos0]] : 0.0f;
        sxr1 = (pos1 >= 0 && pos1 < g.p) ? sx[perm[pos1]] : 0.0f;
        asm volatile("bar.sync 1, 128;" ::: "memory");  // previous block's reads of sxc are done
        S->sxc[row] = sxr0;
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
#pragma unroll
      for (int c = 0; c < 4 * NBAND; c++) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * kAccCols + (uint32_t)c * 32u;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (FP8) {  // accumulator = exact integer * 2^-18 in fp32
#pragma unroll
          for (int q = 0; q < 32; q++) {
            const float f = __uint_as_float(v[q]) * 262144.0f;
            v[q] = out_f32 ? __float_as_uint(f) : (uint32_t)__float2int_rn(f);
          }
        } else if (out_f32) {
#pragma unroll
          for (int q = 0; q < 32; q++) v[q] = __float_as_uint(__int2float_rn((int)v[q]));
        }
        if (sx && out_f32) {
          const float sr = (c < 4 ? sxr0 : sxr1) * inv_n;
#pragma unroll
          for (int q = 0; q < 32; q++) v[q] = __float_as_uint(fmaf(-sr, S->sxc[(c & 3) * 32 + q], __uint_as_float(v[q])));
        }
        if (ok) {
#pragma unroll
          for (int q = 0; q < 32; q += 4)
            *reinterpret_cast<uint4*>(out + c * 32 + q) = make_uint4(v[q], v[q + 1], v[q + 2], v[q + 3]);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&S->tmem_empty[as]);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2u * kAccCols) : "memory");
  }
}

// Debug / test cross-check: the same Gram blocks with dp4a on the CUDA cores.
__global__ void __launch_bounds__(256) gram_simt_kernel(GenoView g, const int* __restrict__ perm, int32_t* __restrict__ gram, int out_f32) {
  const int blk = blockIdx.x;
  const int nw = (int)(g.ld >> 2);
  for (int idx = threadIdx.x; idx < 128 * 128; idx += blockDim.x) {
    const int i = idx >> 7, j = idx & 127;
    const int pi = blk * 128 + i, pj = blk * 128 + j;
    int acc = 0;
    if (pi < g.p && pj < g.p) {
      const int* ci = reinterpret_cast<const int*>(g.x8 + (int64_t)perm[pi] * g.ld);
      const int* cj = reinterpret_cast<const int*>(g.x8 + (int64_t)perm[pj] * g.ld);
      for (int w = 0; w < nw; w++) acc = __dp4a(ci[w], cj[w], acc);
    }
    gram[(size_t)blk * 128 * 128 + idx] = out_f32 ? __float_as_int(__int2float_rn(acc)) : acc;
  }
}

}  // namespace

template <int NBAND, bool FP8>
static void launch_gram_band(const GenoView& g, const int* perm, int nblocks, void* gram, int out_f32, int* err, int num_sms,
                             const float* sx, const void* tmap, cudaStream_t st) {
  const size_t smem = (size_t)GramCfg<NBAND>::kStages * GramCfg<NBAND>::kSub * NBAND * kTileBytes + sizeof(GramSmem<NBAND>) + 1024;
  const int grid = nblocks < num_sms ? nblocks : num_sms;
  const char* de = getenv("BWGR_GRAM_DBG");
  const int dbg = de ? atoi(de) : 0;
  CUtensorMap tm;
  memset(&tm, 0, sizeof tm);
  if (tmap) memcpy(&tm, tmap, sizeof tm);
  const float inv_n = 1.0f / (float)g.n;
  int32_t* out = static_cast<int32_t*>(gram);
#define BWGR_GRAM_LAUNCH(PROD)                                                                                              \
  do {                                                                                                                      \
    cudaFuncSetAttribute(gram_tc_kernel<NBAND, FP8, PROD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
    gram_tc_kernel<NBAND, FP8, PROD><<<grid, 416, smem, st>>>(g, perm, nblocks, out, out_f32, err, sx, inv_n, dbg, tm);     \
  } while (0)
  if (FP8 && g.x2 && !tmap) BWGR_GRAM_LAUNCH(2);   // packed 2-bit shadow copy: codes 0..2, exact E4M3 products
  else if (tmap) BWGR_GRAM_LAUNCH(1);
  else BWGR_GRAM_LAUNCH(0);
#undef BWGR_GRAM_LAUNCH
}
void launch_gram_tc(const GenoView& g, const int* perm, int nblocks, void* gram, int out_f32, int nband, int fp8_codes,
                    int* err, int num_sms, const float* sx, const void* tmap, cudaStream_t st) {
  if (fp8_codes) {
    if (nband == 2) launch_gram_band<2, true>(g, perm, nblocks, gram, out_f32, err, num_sms, sx, tmap, st);
    else launch_gram_band<1, true>(g, perm, nblocks, gram, out_f32, err, num_sms, sx, tmap, st);
  } else {
    if (nband == 2) launch_gram_band<2, false>(g, perm, nblocks, gram, out_f32, err, num_sms, sx, tmap, st);
    else launch_gram_band<1, false>(g, perm, nblocks, gram, out_f32, err, num_sms, sx, tmap, st);
  }
}

// Tensor map of the int8 genotype store as a 2-D tensor [p columns][ld rows], box = 128 rows x 1 column, SWIZZLE_128B:
// the descriptor tile::gather4 needs.  Built through the driver entry point (no link-time dependency on libcuda).
// tmap_out: 128 bytes, 64-byte aligned.  Returns false if the driver cannot provide it (the kernels then gather with cp.async).
bool make_geno_tensor_map(const int8_t* x8, int64_t ld, int64_t p, void* tmap_out) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
      qres != cudaDriverEntryPointSuccess)
    return false;
  const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)p};
  const cuuint64_t strides[1] = {(cuuint64_t)ld};
  const cuuint32_t box[2] = {128, 1}, estr[2] = {1, 1};
  CUtensorMap tm;
  const CUresult r = reinterpret_cast<EncodeFn>(fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<int8_t*>(x8), dims, strides, box,
                                                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  memcpy(tmap_out, &tm, sizeof tm);
  return true;
}
void launch_gram_simt(const GenoView& g, const int* perm, int nblocks, void* gram, int out_f32, cudaStream_t st) {
  gram_simt_kernel<<<nblocks, 256, 0, st>>>(g, perm, static_cast<int32_t*>(gram), out_f32);
}

}  // namespace bwgr
