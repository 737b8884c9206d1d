// block_inv.cu -- explicit inverse of the in-block system of the linear rules, one 128 x 128 block per CTA.
//
// Within a block of 128 markers the Gauss-Seidel steps of a linear rule (emRR :335, emBA :107-111, BayesRR :835,
// BayesA :615) are the unit-lower-triangular system (I + A L) dE = A g + c, L = strictly lower part of X_B'X_B,
// A = diag(a_i), a_i = kappa / (xx_i + lambda_i).  a_i does not depend on the residuals, so T = (I + A L)^-1 can be
// formed for EVERY block of the sweep before the sweep starts, massively parallel and off the sweep's critical path;
// the solver CTA of the pipelined sweep then applies it as one lower-triangular mat-vec on four warps instead of walking
// four dependent 32-marker steps (sweep_pipe.cu).
//
// One CTA per block, one warp per block column of T (32 columns), lane = column.  A lane walks its column down the 32-row
// tiles: the running sums of a tile (32) live in registers; the finished part of the column is parked in shared memory (it is the
// multiplier of the later tile products); the Gram tiles are staged in shared memory and read as warp-wide broadcasts, off-diagonal tiles as their
// mirror image so that a 128-bit load runs down the 32 accumulators.  The loops over tiles stay rolled: fully unrolled, this kernel spent six of every
// seven issue slots waiting for instructions (ncu: stall no_instruction 5.9 per issue).
#include "kernels.h"

namespace bwgr {

namespace {

constexpr int kTS = 36;           // row stride (floats) of a 32 x 32 tile in shared memory
constexpr int kTileF = 32 * kTS;
constexpr int kXsF = 32 * (96 + 64 + 32);  // parked columns of T: column tile j keeps rows 32 j .. 95

__device__ __forceinline__ int tri(int hi, int lo) { return hi * (hi + 1) / 2 + lo; }

// kappa = 2 for emBA (the reference applies the residual update twice); the penalty of a marker follows marker_lambda<>
// of common.cuh (same float expressions, so the inverse and the right-hand side of the solve use the same a_i)
__global__ void __launch_bounds__(128) block_inverse_kernel(const int* __restrict__ perm, int p, const float* __restrict__ gram,
                                                            int gstride, const float* __restrict__ xx,
                                                            const float* __restrict__ vbv, const SysScalars* __restrict__ sc,
                                                            float kappa, int model, float* __restrict__ tinv) {
  extern __shared__ float sm[];
  float* xs = sm;                        // rows 32 j .. 95 of the 32 columns of column tile j, [row][lane]: 6144 floats in all
  float* av = sm + kXsF;                 // [128]
  float* Gs = sm + kXsF + 128;           // 10 lower-triangle tiles (hi, lo), row stride kTS; off-diagonal tiles transposed
  const int blk = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const int j = ((tid >> 5) + blk) & 3;  // the heavy column tile (j = 0: ten tiles) sits on a different scheduler in neighbouring CTAs
  const int col = 32 * j + lane;
  {
    const SysScalars s = sc[0];
    const int pos = blk * 128 + tid;
    float a = 0.0f;
    if (pos < p) {
      const int m = perm[pos];
      float lmb;
      switch (model) {
        case M_EMBA: case M_BA: lmb = s.ve * (1.0f / vbv[m]); break;
        case M_BL: lmb = s.sweep == 0 ? s.ve * (1.0f / vbv[m]) : sqrtf(s.Rho * s.ve / vbv[m]); break;
        case M_EMDE: lmb = vbv[m]; break;
        default: lmb = s.lmb; break;
      }
      a = kappa / (xx[m] + lmb);
    }
    av[tid] = a;
  }
  {
    // rows = markers of the block, columns 0..127 = the (symmetric) diagonal block.  Off-diagonal tiles are kept transposed,
    // Gs[q][r] = G[32 hi + r][32 lo + q], read as the mirror tile (lo, hi): the product loop walks q with a 128-bit broadcast down r.
    // (Reading the band directly through L1 was tried: with this much shared memory per SM the L1 left over thrashes, 135 us.)
    const float* G = gram + (size_t)blk * 128 * gstride;
    // 20 x 16 bytes per thread, ten loads in flight at a time (one after the other they are twenty L2 round trips)
#pragma unroll
    for (int half = 0; half < 2; half++) {
      float4 v[10];
#pragma unroll
      for (int u = 0; u < 10; u++) {
        const int idx = tid + 128 * (10 * half + u);
        const int tile = idx >> 8, rr = (idx >> 3) & 31, c4 = idx & 7;
        const int hi = tile >= 6 ? 3 : tile >= 3 ? 2 : tile >= 1 ? 1 : 0, lo = tile - tri(hi, 0);
        const int srow = hi == lo ? 32 * hi + rr : 32 * lo + rr, scol = hi == lo ? 32 * lo : 32 * hi;
        v[u] = __ldg(reinterpret_cast<const float4*>(G + (size_t)srow * gstride + scol + 4 * c4));
      }
#pragma unroll
      for (int u = 0; u < 10; u++) {
        const int idx = tid + 128 * (10 * half + u);
        const int tile = idx >> 8, rr = (idx >> 3) & 31, c4 = idx & 7;
        *reinterpret_cast<float4*>(Gs + (size_t)tile * kTileF + rr * kTS + 4 * c4) = v[u];
      }
    }
  }
  __syncthreads();
  float* out = tinv + (size_t)blk * 128 * 128 + col;
  float* xw = xs + (j == 0 ? 0 : j == 1 ? 3072 : 5120) + lane;  // this warp's part of xs: rows 32 j .. 95, [row - 32 j][lane]
#pragma unroll 1
  for (int i = j; i < 4; i++) {
    float acc[32];
#pragma unroll
    for (int r = 0; r < 32; r++) acc[r] = 0.0f;
    // contributions of the tiles already known: acc_r += G[32 i + r][32 k + q] x[32 k + q], k = j .. i-1
#pragma unroll 1
    for (int k = j; k < i; k++) {
      const float* gq = Gs + (size_t)tri(i, k) * kTileF;  // transposed tile: gq[q][r]
      const float* xq = xw + (32 * (k - j)) * 32;
#pragma unroll 4
      for (int q = 0; q < 32; q++) {
        const float x = xq[q * 32];
#pragma unroll
        for (int r4 = 0; r4 < 8; r4++) {
          const float4 g4 = *reinterpret_cast<const float4*>(gq + q * kTS + 4 * r4);
          acc[4 * r4 + 0] = fmaf(g4.x, x, acc[4 * r4 + 0]); acc[4 * r4 + 1] = fmaf(g4.y, x, acc[4 * r4 + 1]);
          acc[4 * r4 + 2] = fmaf(g4.z, x, acc[4 * r4 + 2]); acc[4 * r4 + 3] = fmaf(g4.w, x, acc[4 * r4 + 3]);
        }
      }
    }
    // the diagonal tile: right-looking substitution, x_r = delta - a_r acc_r, then acc_r2 += G_ii[r2][r] x_r for r2 > r
    const float* gd = Gs + (size_t)tri(i, i) * kTileF;
    const float* ai = av + 32 * i;
#pragma unroll
    for (int r = 0; r < 32; r++) {
      const float x = ((i == j && r == lane) ? 1.0f : 0.0f) - ai[r] * acc[r];
      if (i < 3) xw[(32 * (i - j) + r) * 32] = x;
      out[(size_t)(32 * i + r) * 128] = x;  // T[32 i + r][32 j + lane]: 128 B per warp store
#pragma unroll
      for (int q4 = (r + 1) / 4; q4 < 8; q4++) {
        const float4 g4 = *reinterpret_cast<const float4*>(gd + r * kTS + 4 * q4);  // row r = column r (symmetric)
        if (4 * q4 + 0 > r) acc[4 * q4 + 0] = fmaf(g4.x, x, acc[4 * q4 + 0]);
        if (4 * q4 + 1 > r) acc[4 * q4 + 1] = fmaf(g4.y, x, acc[4 * q4 + 1]);
        if (4 * q4 + 2 > r) acc[4 * q4 + 2] = fmaf(g4.z, x, acc[4 * q4 + 2]);
        if (4 * q4 + 3 > r) acc[4 * q4 + 3] = fmaf(g4.w, x, acc[4 * q4 + 3]);
      }
    }
    __syncwarp();
  }
}

}  // namespace

void launch_block_inverse(int model, const int* perm, int p, int nblocks, const float* gram, int nband, const float* xx,
                          const float* vbv, const SysScalars* sc, float* tinv, cudaStream_t st) {
  const size_t smem = (size_t)(kXsF + 128 + 10 * kTileF) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(block_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  const float kappa = model == M_EMBA ? 2.0f : 1.0f;
  block_inverse_kernel<<<nblocks, 128, smem, st>>>(perm, p, gram, nband * 128, xx, vbv, sc, kappa, rule_model(model), tinv);
}

}  // namespace bwgr
