// mrr.cu -- device helpers of the multivariate ridge (MRR3 / MRR3F, RcppEigen20230423.cpp:318-701, :704-1079).
//
// The per-marker k x k solve of the reference (LHS = iG + xx_J diag(iVe), :504-516) is diagonalised once per sweep:
// with S = diag(iVe)^(1/2) and S^-1 iG S^-1 = U Lambda U', the rotated effects b~ = U'S b and residuals E~ = E S U obey
// k INDEPENDENT ridge recurrences  b~_t <- (x'e~_t + xx b~_t) / (xx + Lambda_t)  -- the blocked sweep kernel with k
// systems (exact up to reassociation; valid for complete Y, the fast path built here).  These kernels do the O((n+p) k^2)
// work around it: the rotations, X'Y (tilde, :420) and the k x k / k-vector reductions of the sweep epilogue (:536-556).
#include "kernels.h"

namespace bwgr {

namespace {

// tilde[t][j] = sum_i x_ij * Y[t][i]  (one CTA per marker; columns are NOT centred here: Y is centred, so X_c'Y = X'Y)
__global__ void __launch_bounds__(256) xty_kernel(GenoView g, const float* __restrict__ Y, int k, float* __restrict__ out) {
  __shared__ float red[8][32];
  const int j = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int8_t* col = g.x8 + (int64_t)j * g.ld;
  float acc[32];
#pragma unroll
  for (int t = 0; t < 32; t++) acc[t] = 0.0f;
  for (int i = tid; i < g.n; i += 256) {
    const float x = (float)col[i];
#pragma unroll
    for (int t = 0; t < 32; t++)
      if (t < k) acc[t] = fmaf(x, Y[(size_t)t * g.ld + i], acc[t]);
  }
#pragma unroll
  for (int t = 0; t < 32; t++) {
    if (t < k) {
      const float v = warp_sum(acc[t]);
      if (lane == 0) red[warp][t] = v;
    }
  }
  __syncthreads();
  if (tid < k) {
    float s = 0.0f;
    for (int w = 0; w < 8; w++) s += red[w][tid];
    out[(size_t)tid * g.p + j] = s;
  }
}

// out[t][i] = sum_s in[s][i] * T[s][t] + sum_s shift[s] * T[s][t]   (i < len; matrices are [k][ld]); amax[t] = max_i |out[t][i]|
__global__ void __launch_bounds__(256) rotate_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t ld, int len,
                                                     int k, const float* __restrict__ T, const float* __restrict__ shift,
                                                     float* __restrict__ amax) {
  __shared__ float Ts[32 * 32], add[32], mx[32];
  const int tid = threadIdx.x;
  for (int q = tid; q < k * k; q += 256) Ts[q] = T[q];
  if (tid < 32) mx[tid] = 0.0f;
  __syncthreads();
  if (tid < k) {
    float a = 0.0f;
    if (shift) for (int s = 0; s < k; s++) a = fmaf(shift[s], Ts[s * k + tid], a);
    add[tid] = a;
  }
  __syncthreads();
  float lmax[32];
#pragma unroll
  for (int t = 0; t < 32; t++) lmax[t] = 0.0f;
  for (int i = blockIdx.x * 256 + tid; i < len; i += gridDim.x * 256) {
    float v[32];
#pragma unroll
    for (int s = 0; s < 32; s++) v[s] = s < k ? in[(size_t)s * ld + i] : 0.0f;
#pragma unroll
    for (int t = 0; t < 32; t++) {
      if (t < k) {
        float a = add[t];
#pragma unroll
        for (int s = 0; s < 32; s++)
          if (s < k) a = fmaf(v[s], Ts[s * k + t], a);
        out[(size_t)t * ld + i] = a;
        lmax[t] = fmaxf(lmax[t], fabsf(a));
      }
    }
  }
  if (amax) {
#pragma unroll
    for (int t = 0; t < 32; t++) {
      if (t < k) {
        float m = lmax[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((tid & 31) == 0) atomicMax(reinterpret_cast<int*>(&mx[t]), __float_as_int(m));  // non-negative floats order as ints
      }
    }
    __syncthreads();
    if (tid < k) atomicMax(reinterpret_cast<int*>(&amax[tid]), __float_as_int(mx[tid]));
  }
}

// mode 0: out[t1*k + t2] = sum_i A[t1][i] * B[t2][i];  mode 1: out[t1] = sum_i (A[t1][i] - B[t1][i])^2;
// mode 2: out[t1] = sum_i A[t1][i]                      (double accumulation, one CTA per output)
__global__ void __launch_bounds__(256) pair_reduce_kernel(const float* __restrict__ A, const float* __restrict__ B, int64_t lda,
                                                          int64_t ldb, int len, int k, int mode, double* __restrict__ out) {
  __shared__ double sh[8];
  const int t1 = blockIdx.x, t2 = mode == 0 ? blockIdx.y : blockIdx.x, tid = threadIdx.x;
  const float* a = A + (size_t)t1 * lda;
  const float* b = B ? B + (size_t)t2 * ldb : nullptr;
  double s = 0.0;
  for (int i = tid; i < len; i += 256) {
    if (mode == 0) s += (double)a[i] * (double)b[i];
    else if (mode == 1) { const double d = (double)a[i] - (double)b[i]; s += d * d; }
    else s += (double)a[i];
  }
  s = warp_sum(s);
  if ((tid & 31) == 0) sh[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    double r = 0.0;
    for (int w = 0; w < 8; w++) r += sh[w];
    out[mode == 0 ? t1 * k + t2 : t1] = r;
  }
}

}  // namespace

void launch_xty(const GenoView& g, const float* Y, int k, float* out, cudaStream_t st) { xty_kernel<<<g.p, 256, 0, st>>>(g, Y, k, out); }

void launch_rotate(const float* in, float* out, int64_t ld, int len, int k, const float* T, const float* shift, float* amax,
                   int num_sms, cudaStream_t st) {
  rotate_kernel<<<num_sms * 2, 256, 0, st>>>(in, out, ld, len, k, T, shift, amax);
}

void launch_pair_reduce(const float* A, const float* B, int64_t lda, int64_t ldb, int len, int k, int mode, double* out,
                        cudaStream_t st) {
  dim3 grid(k, mode == 0 ? k : 1);
  pair_reduce_kernel<<<grid, 256, 0, st>>>(A, B, lda, ldb, len, k, mode, out);
}

}  // namespace bwgr
