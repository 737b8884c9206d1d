// sweep_blocked.cu -- the blocked exact Gauss-Seidel / Gibbs sweep (K1 + K5 + K2 of SURVEY 2c).
//
// One persistent cooperative kernel per sweep.  CTA c owns the row slab [c*R, (c+1)*R) of every
// genotype column and the matching slab of the residuals E (kept in shared memory for the sweep).
// For each block of 128 markers (in this sweep's order):
//   1. the slab of X_B is staged in shared memory once (cp.async, double buffered, prefetched a block
//      ahead -- it does not depend on E);
//   2. partial g_B = X_B' E over the slab, converted to 64-bit fixed point and added to the block
//      accumulator in L2 (integer atomics: the grid-wide sum is exact and order independent);
//   3. one grid barrier; every CTA then reads the same g_B and redundantly runs the sequential
//      in-block solve on the precomputed Gram block X_B'X_B (gram_tc.cu) held in shared memory:
//      marker jj uses g_jj - sum_{i<jj} G[jj][i]*de_i, i.e. exactly the reference's Gauss-Seidel order
//      (Rcpp20260726ai.cpp:332-337) up to float reassociation;
//   4. E_slab -= X_B_slab * dE_B from the same staged tile: X is read from HBM once per sweep.
// The solve is replicated on all CTAs instead of broadcast, which saves the second grid barrier.
#include <cooperative_groups.h>

#include "kernels.h"

namespace bwgr {

namespace {

constexpr int kThreads = 256;
constexpr int kMaxSysGroup = 4;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__host__ __device__ inline int padded_rows(int R) { return ((R >> 4) & 1) ? R : R + 16; }

struct MarkerIn { float b0, xx, vbj, pad; };

template <int MODEL>
__global__ void __launch_bounds__(kThreads, 1) sweep_blocked_kernel(SweepArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int R = a.rows_per_cta, RP = padded_rows(R);
  const int row0 = blockIdx.x * R;
  const int ns = a.nsys, p = a.g.p;
  const int G = gridDim.x;
  const int nchunk = R >> 4;

  // ---- shared memory carve-up
  unsigned char* Xs = smem;                                               // [2][128][RP]
  float* Gs = reinterpret_cast<float*>(Xs + 2 * 128 * RP);                // [128][128]
  float* Es = Gs + 128 * 128;                                             // [ns][R]
  float* gpart = Es + ns * R;                                             // [2][ns][128]
  float* dlt = gpart + 2 * ns * 128;                                      // [ns][128]
  MarkerIn* mk = reinterpret_cast<MarkerIn*>(dlt + ns * 128);             // [ns][128]
  MarkerDraws* drw = reinterpret_cast<MarkerDraws*>(mk + ns * 128);       // [ns][128]
  float* upd = reinterpret_cast<float*>(drw + (model_is_gibbs(MODEL) ? ns * 128 : 0));  // [kMaxSysGroup][kThreads] float4 partial updates
  __shared__ SysScalars sc[32];
  __shared__ int s_fail;

  if (tid < ns) sc[tid] = a.sc[tid];
  if (tid == 0) s_fail = 0;
  for (int s = 0; s < ns; s++)
    for (int i = tid; i < R; i += kThreads) {
      const int r = row0 + i;
      Es[s * R + i] = (r < a.g.ld) ? a.e[(size_t)s * a.g.ld + r] : 0.0f;
    }

  auto issue_tile = [&](int blk) {
    if (blk < a.nblocks) {
      unsigned char* dst = Xs + (blk & 1) * 128 * RP;
      const int total = 128 * nchunk;
      for (int idx = tid; idx < total; idx += kThreads) {
        const int m = idx / nchunk, c = idx - m * nchunk;
        const int pos = blk * 128 + m;
        const int r = row0 + 16 * c;
        if (pos < p && r < a.g.ld) {
          cp_async16(smem_u32(dst + m * RP + 16 * c), a.g.x8 + (int64_t)a.perm[pos] * a.g.ld + r);
        } else {
          *reinterpret_cast<uint4*>(dst + m * RP + 16 * c) = make_uint4(0, 0, 0, 0);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto issue_gram = [&](int blk) {
    const float* src = a.gram + (size_t)blk * 128 * 128;
    for (int idx = tid; idx < 128 * 128 / 4; idx += kThreads) cp_async16(smem_u32(Gs + 4 * idx), src + 4 * idx);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  issue_tile(0);
  __syncthreads();

  const float inv_q = 1.0f / a.g_quantum;
  bool fail = false;

#pragma unroll 1
  for (int blk = 0; blk < a.nblocks; blk++) {
    const unsigned char* Xt = Xs + (blk & 1) * 128 * RP;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // ---- 1. partial g over the slab: thread = (column, half of the row chunks)
    {
      const int col = tid & 127, half = tid >> 7;
      for (int s0 = 0; s0 < ns; s0 += kMaxSysGroup) {
        float acc[kMaxSysGroup];
#pragma unroll
        for (int q = 0; q < kMaxSysGroup; q++) acc[q] = 0.0f;
        for (int c = half; c < nchunk; c += 2) {
          const uint4 w = *reinterpret_cast<const uint4*>(Xt + col * RP + 16 * c);
          const uint32_t ww[4] = {w.x ^ 0x80808080u, w.y ^ 0x80808080u, w.z ^ 0x80808080u, w.w ^ 0x80808080u};
          float xf[16];
#pragma unroll
          for (int q = 0; q < 16; q++) xf[q] = byte_to_float(ww[q >> 2], q & 3);
#pragma unroll
          for (int s = 0; s < kMaxSysGroup; s++) {
            if (s0 + s < ns) {
              const float4* ev = reinterpret_cast<const float4*>(Es + (s0 + s) * R + 16 * c);
#pragma unroll
              for (int q = 0; q < 4; q++) {
                const float4 e4 = ev[q];
                acc[s] = fmaf(xf[4 * q + 0], e4.x, acc[s]);
                acc[s] = fmaf(xf[4 * q + 1], e4.y, acc[s]);
                acc[s] = fmaf(xf[4 * q + 2], e4.z, acc[s]);
                acc[s] = fmaf(xf[4 * q + 3], e4.w, acc[s]);
              }
            }
          }
        }
#pragma unroll
        for (int s = 0; s < kMaxSysGroup; s++)
          if (s0 + s < ns) gpart[(half * ns + s0 + s) * 128 + col] = acc[s];
      }
    }
    __syncthreads();
    // combine the two halves, quantise, add to the block accumulator (exact integer sum in L2)
    for (int idx = tid; idx < ns * 128; idx += kThreads) {
      const float v = gpart[idx] + gpart[ns * 128 + idx];
      if (!(fabsf(v) <= a.g_limit)) fail = true;
      const long long q = __float2ll_rn(v * inv_q);
      atomicAdd(reinterpret_cast<unsigned long long*>(a.gacc) + (size_t)(blk % 3) * ns * 128 + idx, (unsigned long long)q);
    }
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      atomicAdd(a.bar, 1u);
    }
    // ---- prefetch while waiting: Gram block of this block, genotype tile of the next
    issue_gram(blk);
    issue_tile(blk + 1);
    // per-marker inputs of this block (do not depend on the barrier)
    for (int idx = tid; idx < ns * 128; idx += kThreads) {
      const int s = idx >> 7, m = idx & 127, pos = blk * 128 + m;
      MarkerIn in = {0.0f, 1.0f, 1.0f, 0.0f};
      if (pos < p && !sc[s].done) {
        const int j = a.perm[pos];
        in.b0 = a.b[(size_t)s * p + j];
        in.xx = a.xx[j];
        in.vbj = (model_has_vbj(MODEL) && a.vbv) ? a.vbv[(size_t)s * p + j] : 1.0f;
        if (model_is_gibbs(MODEL))
          drw[idx] = marker_draws(MODEL, (uint32_t)j, (uint32_t)sc[s].sweep, (uint32_t)(a.chain0 + s), sc[s].df, a.seed_lo, a.seed_hi);
      }
      mk[idx] = in;
    }
    // ---- 2. grid barrier (monotonic counter; bounded spin so a bug cannot hang the GPU)
    if (tid == 0) {
      const unsigned int target = (unsigned int)(blk + 1) * (unsigned int)G;
      unsigned int spins = 0;
      while (true) {
        unsigned int v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.bar) : "memory");
        if (v >= target) break;
        if (++spins > (1u << 24)) { s_fail = 1; atomicExch(a.err, 3); break; }
      }
    }
    asm volatile("cp.async.wait_group 1;" ::: "memory");  // Gram block landed (next tile may still fly)
    __syncthreads();

    // ---- 3. sequential in-block solve, one warp per system, replicated on every CTA
    for (int s = warp; s < ns; s += kThreads / 32) {
      if (sc[s].done) {
        for (int t = 0; t < 4; t++) dlt[s * 128 + 32 * t + lane] = 0.0f;
        continue;
      }
      const SysScalars S = sc[s];
      float g[4], nb[4], nd[4], nv[4], de[4];
      const long long* gq = a.gacc + (size_t)(blk % 3) * ns * 128 + s * 128;
#pragma unroll
      for (int t = 0; t < 4; t++) {
        long long q;
        asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(q) : "l"(gq + 32 * t + lane) : "memory");
        g[t] = (float)((double)q * (double)a.g_quantum);
        nb[t] = 0.0f; nd[t] = 1.0f; nv[t] = 1.0f; de[t] = 0.0f;
      }
      const int nvalid = min(128, p - blk * 128);
#pragma unroll
      for (int t = 0; t < 4; t++) {
#pragma unroll 8
        for (int i = 0; i < 32; i++) {
          const int jj = 32 * t + i;
          if (jj >= nvalid) break;
          const float gc = __shfl_sync(0xffffffffu, g[t], i);
          const MarkerIn in = mk[s * 128 + jj];
          MarkerDraws dr;
          if (model_is_gibbs(MODEL)) dr = drw[s * 128 + jj];
          else { dr.z1 = dr.z2 = dr.u = 0.0f; dr.chi = 1.0f; }
          const RuleOut r = marker_rule<MODEL>(gc, in.xx, in.b0, in.vbj, S, dr);
          if (lane == i) { nb[t] = r.b; nd[t] = r.d; nv[t] = r.vbj; de[t] = r.de; }
          const float* grow = Gs + jj * 128 + lane;
#pragma unroll
          for (int tt = 0; tt < 4; tt++)
            if (tt >= t) g[tt] = fmaf(-grow[32 * tt], r.de, g[tt]);
        }
      }
#pragma unroll
      for (int t = 0; t < 4; t++) {
        const int jj = 32 * t + lane;
        dlt[s * 128 + jj] = (jj < nvalid) ? de[t] : 0.0f;
        if (blockIdx.x == 0 && jj < nvalid) {
          const int j = a.perm[blk * 128 + jj];
          a.b[(size_t)s * p + j] = nb[t];
          if (model_has_d(MODEL) && a.d) a.d[(size_t)s * p + j] = nd[t];
          if (model_has_vbj(MODEL) && MODEL != M_KMUP && a.vbv) a.vbv[(size_t)s * p + j] = nv[t];
        }
      }
    }
    // recycle the accumulator of block blk+2 (its last readers passed this block's barrier)
    if (blockIdx.x == 0 && blk + 2 < a.nblocks)
      for (int idx = tid; idx < ns * 128; idx += kThreads) a.gacc[(size_t)((blk + 2) % 3) * ns * 128 + idx] = 0;
    __syncthreads();

    // ---- 4. E_slab -= X_B_slab * dE_B: thread = (row quad, column group); partials combined in smem
    {
      const int nq = R >> 2;                       // row quads in the slab
      const int ngrp = kThreads / nq > 0 ? min(kThreads / nq, 8) : 1;
      const int q = tid % nq, cg = tid / nq;
      for (int s0 = 0; s0 < ns; s0 += kMaxSysGroup) {
        float acc[kMaxSysGroup][4];
#pragma unroll
        for (int s = 0; s < kMaxSysGroup; s++) acc[s][0] = acc[s][1] = acc[s][2] = acc[s][3] = 0.0f;
        if (cg < ngrp) {
          for (int qq = q; qq < nq; qq += kThreads) {  // nq > kThreads: single group, strided quads
            for (int col = cg; col < 128; col += ngrp) {
              const uint32_t w = *reinterpret_cast<const uint32_t*>(Xt + col * RP + 4 * qq) ^ 0x80808080u;
              const float x0 = byte_to_float(w, 0), x1 = byte_to_float(w, 1), x2 = byte_to_float(w, 2), x3 = byte_to_float(w, 3);
#pragma unroll
              for (int s = 0; s < kMaxSysGroup; s++) {
                if (s0 + s < ns) {
                  const float dv = dlt[(s0 + s) * 128 + col];
                  acc[s][0] = fmaf(x0, dv, acc[s][0]); acc[s][1] = fmaf(x1, dv, acc[s][1]);
                  acc[s][2] = fmaf(x2, dv, acc[s][2]); acc[s][3] = fmaf(x3, dv, acc[s][3]);
                }
              }
            }
            if (nq > kThreads) {  // no column split possible: apply directly
#pragma unroll
              for (int s = 0; s < kMaxSysGroup; s++)
                if (s0 + s < ns) {
                  float4* ev = reinterpret_cast<float4*>(Es + (s0 + s) * R + 4 * qq);
                  float4 e4 = *ev;
                  e4.x -= acc[s][0]; e4.y -= acc[s][1]; e4.z -= acc[s][2]; e4.w -= acc[s][3];
                  *ev = e4;
                  acc[s][0] = acc[s][1] = acc[s][2] = acc[s][3] = 0.0f;
                }
            }
          }
        }
        if (nq <= kThreads) {
#pragma unroll
          for (int s = 0; s < kMaxSysGroup; s++)
            if (s0 + s < ns && cg < ngrp)
              *reinterpret_cast<float4*>(upd + ((size_t)s * kThreads + cg * nq + q) * 4) = make_float4(acc[s][0], acc[s][1], acc[s][2], acc[s][3]);
          __syncthreads();
          if (tid < nq) {
#pragma unroll
            for (int s = 0; s < kMaxSysGroup; s++)
              if (s0 + s < ns) {
                float4* ev = reinterpret_cast<float4*>(Es + (s0 + s) * R + 4 * tid);
                float4 e4 = *ev;
                for (int c2 = 0; c2 < ngrp; c2++) {
                  const float4 u = *reinterpret_cast<const float4*>(upd + ((size_t)s * kThreads + c2 * nq + tid) * 4);
                  e4.x -= u.x; e4.y -= u.y; e4.z -= u.z; e4.w -= u.w;
                }
                *ev = e4;
              }
          }
          __syncthreads();
        }
      }
    }
    if (s_fail) break;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (fail) atomicExch(a.err, 4);
  for (int s = 0; s < ns; s++)
    for (int i = tid; i < R; i += kThreads) {
      const int r = row0 + i;
      if (r < a.g.ld) a.e[(size_t)s * a.g.ld + r] = Es[s * R + i];
    }
}

size_t smem_bytes(int R, int ns, bool gibbs) {
  const int RP = padded_rows(R);
  return (size_t)2 * 128 * RP + 128 * 128 * 4 + (size_t)ns * R * 4 + (size_t)2 * ns * 128 * 4 + (size_t)ns * 128 * 4 +
         (size_t)ns * 128 * sizeof(MarkerIn) + (gibbs ? (size_t)ns * 128 * sizeof(MarkerDraws) : 0) +
         (size_t)kMaxSysGroup * kThreads * 16 + 64;
}

template <int MODEL>
void launch_model(const SweepArgs& a, int grid, cudaStream_t st) {
  const size_t smem = smem_bytes(a.rows_per_cta, a.nsys, model_is_gibbs(MODEL));
  cudaFuncSetAttribute(sweep_blocked_kernel<MODEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  SweepArgs args = a;
  void* params[] = {&args};
  cudaLaunchCooperativeKernel((void*)sweep_blocked_kernel<MODEL>, dim3(grid), dim3(kThreads), params, smem, st);
}

}  // namespace

size_t sweep_blocked_smem(int rows_per_cta, int nsys) { return smem_bytes(rows_per_cta, nsys, true); }

void launch_sweep_blocked(const SweepArgs& a, int grid, cudaStream_t st) {
  switch (a.model) {
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
    default: break;
  }
}

}  // namespace bwgr
