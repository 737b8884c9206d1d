// block_inv.cu -- explicit inverse of the in-block system of the linear rules, one 128 x 128 block per CTA.
//
// Within a block of 128 markers the Gauss-Seidel steps of a linear rule (emRR :335, emBA :107-111, BayesRR :835,
// BayesA :615) are the unit-lower-triangular system (I + A L) dE = A g + c, L = strictly lower part of X_B'X_B,
// A = diag(a_i), a_i = kappa / (xx_i + lambda_i).  a_i does not depend on the residuals, so T = (I + A L)^-1 can be
// formed for EVERY block of the sweep before the sweep starts, massively parallel and off the sweep's critical path;
// the solver CTA of the pipelined sweep then applies it as one lower-triangular mat-vec on four warps instead of walking
// four dependent 32-marker steps (sweep_pipe.cu).
//
// One CTA per block, warp j = block column j of T (32 columns), lane = column.  A lane walks its column down the four
// 32-row tiles: the running sums and the column live in registers, the Gram tiles are read from shared memory as warp-wide
// broadcasts (every lane needs the same G element), so the work is ~8 k FMAs per lane at one FMA per cycle.
#include "kernels.h"

namespace bwgr {

namespace {

constexpr int kTS = 36;           // row stride (floats) of a 32 x 32 tile in shared memory
constexpr int kTileF = 32 * kTS;

__device__ __forceinline__ int tri(int hi, int lo) { return hi * (hi + 1) / 2 + lo; }

// kappa = 2 for emBA (the reference applies the residual update twice); the penalty of a marker follows marker_lambda<>
// of common.cuh (same float expressions, so the inverse and the right-hand side of the solve use the same a_i)
__global__ void __launch_bounds__(128) block_inverse_kernel(const int* __restrict__ perm, int p, const float* __restrict__ gram,
                                                            int gstride, const float* __restrict__ xx,
                                                            const float* __restrict__ vbv, const SysScalars* __restrict__ sc,
                                                            float kappa, int model, float* __restrict__ tinv) {
  extern __shared__ float sm[];
  float* Gs = sm;                     // 10 lower-triangle tiles (hi, lo): G[32 hi + r][32 lo + q]
  float* av = sm + 10 * kTileF;       // [128]
  const int blk = blockIdx.x, tid = threadIdx.x, lane = tid & 31, j = tid >> 5;
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
    const float* G = gram + (size_t)blk * 128 * gstride;
    for (int idx = tid; idx < 10 * 32 * 8; idx += 128) {
      const int tile = idx >> 8, rr = (idx >> 3) & 31, c4 = idx & 7;
      int hi = 0;
      while (tri(hi + 1, 0) <= tile) hi++;
      const int lo = tile - tri(hi, 0);
      *reinterpret_cast<float4*>(Gs + (size_t)tile * kTileF + rr * kTS + 4 * c4) =
          __ldg(reinterpret_cast<const float4*>(G + (size_t)(32 * hi + rr) * gstride + 32 * lo + 4 * c4));
    }
  }
  __syncthreads();

  float xt[4][32];  // column (j, lane) of T, tile by tile (tiles above the diagonal tile j are not used)
  float* out = tinv + (size_t)blk * 128 * 128 + 32 * j + lane;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    if (i < j) continue;
    float acc[32];
#pragma unroll
    for (int r = 0; r < 32; r++) acc[r] = 0.0f;
    // contributions of the tiles already known: acc += G_ik x_k, k = j .. i-1
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (k >= j && k < i) {
        const float* gt = Gs + (size_t)tri(i, k) * kTileF;
#pragma unroll
        for (int r = 0; r < 32; r++) {
          float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
          for (int q4 = 0; q4 < 8; q4++) {
            const float4 g4 = *reinterpret_cast<const float4*>(gt + r * kTS + 4 * q4);
            s0 = fmaf(g4.x, xt[k][4 * q4 + 0], s0); s1 = fmaf(g4.y, xt[k][4 * q4 + 1], s1);
            s0 = fmaf(g4.z, xt[k][4 * q4 + 2], s0); s1 = fmaf(g4.w, xt[k][4 * q4 + 3], s1);
          }
          acc[r] += s0 + s1;
        }
      }
    }
    // the diagonal tile: right-looking substitution, x_r = delta - a_r acc_r, then acc_r2 += G_ii[r2][r] x_r for r2 > r
    const float* gd = Gs + (size_t)tri(i, i) * kTileF;
#pragma unroll
    for (int r = 0; r < 32; r++) {
      const float x = ((i == j && r == lane) ? 1.0f : 0.0f) - av[32 * i + r] * acc[r];
      xt[i][r] = x;
#pragma unroll
      for (int q4 = (r + 1) / 4; q4 < 8; q4++) {
        const float4 g4 = *reinterpret_cast<const float4*>(gd + r * kTS + 4 * q4);  // row r = column r (symmetric)
        if (4 * q4 + 0 > r) acc[4 * q4 + 0] = fmaf(g4.x, x, acc[4 * q4 + 0]);
        if (4 * q4 + 1 > r) acc[4 * q4 + 1] = fmaf(g4.y, x, acc[4 * q4 + 1]);
        if (4 * q4 + 2 > r) acc[4 * q4 + 2] = fmaf(g4.z, x, acc[4 * q4 + 2]);
        if (4 * q4 + 3 > r) acc[4 * q4 + 3] = fmaf(g4.w, x, acc[4 * q4 + 3]);
      }
    }
#pragma unroll
    for (int r = 0; r < 32; r++) out[(size_t)(32 * i + r) * 128] = xt[i][r];  // T[32 i + r][32 j + lane]: 128 B per warp store
  }
}

}  // namespace

void launch_block_inverse(int model, const int* perm, int p, int nblocks, const float* gram, int nband, const float* xx,
                          const float* vbv, const SysScalars* sc, float* tinv, cudaStream_t st) {
  const size_t smem = (size_t)(10 * kTileF + 128) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(block_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  const float kappa = model == M_EMBA ? 2.0f : 1.0f;
  block_inverse_kernel<<<nblocks, 128, smem, st>>>(perm, p, gram, nband * 128, xx, vbv, sc, kappa, rule_model(model), tinv);
}

}  // namespace bwgr
