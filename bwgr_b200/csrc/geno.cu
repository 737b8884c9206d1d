// geno.cu -- genotype store kernels: packing (f64 -> int8 -> 2-bit), unpack, integer-exact column
// statistics (K4 of SURVEY 2c; replaces xx[j]=squaredNorm / fvar at Rcpp20260726ai.cpp:312-316) and
// the final GEBV pass hat = mu + X b (K10; :346).  All HBM-bound streaming kernels: 16-byte
// vector loads, one column per CTA (stats) or a row tile x column split per CTA (gemv).
#include "kernels.h"

namespace bwgr {

__global__ void pack_f64_kernel(const double* __restrict__ src, int64_t ld_src, int n, int pc, int8_t* __restrict__ dst,
                                int64_t ld, int lo, int hi, int* bad) {
  const int j = blockIdx.x;
  const double* s = src + (int64_t)j * ld_src;
  int8_t* d = dst + (int64_t)j * ld;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < ld; i += gridDim.y * blockDim.x) {
    int8_t v = 0;
    if (i < n) {
      const double x = s[i];
      const double r = rint(x);
      if (!(x == r) || r < (double)lo || r > (double)hi) atomicExch(bad, 1);
      else v = (int8_t)(int)r;
    }
    d[i] = v;
  }
}
void launch_pack_f64(const double* src, int64_t ld_src, int n, int pc, int8_t* dst, int64_t ld, int lo, int hi, int* bad,
                     cudaStream_t st) {
  dim3 grid((unsigned)pc, (unsigned)((ld + 1023) / 1024 > 64 ? 64 : (ld + 1023) / 1024));  // columns on x (p can exceed the 65,535 limit of y)
  pack_f64_kernel<<<grid, 256, 0, st>>>(src, ld_src, n, pc, dst, ld, lo, hi, bad);
}

// PLINK .bed (variant-major): column j = ceil(n/4) bytes, sample i in bits 2*(i%4) of byte i/4: 00 = two copies of allele A1, 10 = one,
// 11 = none, 01 = missing.  Decoded to the additive count of A1 (what `plink --recode A` writes).  missing: 0..2 = that code,
// -2 = the rounded mean of the column's observed codes (integer stand-in for IMP, Rcpp20260726ai.cpp:1316-1335), -1 = an error (*bad).
// nmiss += the number of missing calls.
__global__ void __launch_bounds__(256) decode_bed_kernel(const uint8_t* __restrict__ bed, int64_t bpc, int n, int8_t* __restrict__ dst, int64_t ld,
                                                         int missing, int* bad, unsigned long long* nmiss) {
  __shared__ int s_sum, s_cnt, s_fill;
  const uint8_t* col = bed + (int64_t)blockIdx.x * bpc;
  int8_t* d = dst + (int64_t)blockIdx.x * ld;
  if (threadIdx.x == 0) { s_sum = 0; s_cnt = 0; s_fill = missing >= 0 ? missing : 0; }
  __syncthreads();
  int sum = 0, cnt = 0, miss = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = (col[i >> 2] >> (2 * (i & 3))) & 3;
    if (c == 1) miss++;
    else { sum += c == 0 ? 2 : c == 2 ? 1 : 0; cnt++; }
  }
  if (miss) {
    atomicAdd(nmiss, (unsigned long long)miss);
    if (missing == -1) atomicExch(bad, 1);
  }
  if (missing == -2) {
    atomicAdd(&s_sum, sum); atomicAdd(&s_cnt, cnt);
    __syncthreads();
    if (threadIdx.x == 0) s_fill = s_cnt > 0 ? (int)rintf((float)s_sum / (float)s_cnt) : 0;
    __syncthreads();
  }
  const int fill = s_fill;
  for (int i = threadIdx.x; i < ld; i += blockDim.x) {
    int v = 0;
    if (i < n) {
      const int c = (col[i >> 2] >> (2 * (i & 3))) & 3;
      v = c == 0 ? 2 : c == 2 ? 1 : c == 3 ? 0 : fill;
    }
    d[i] = (int8_t)v;
  }
}
void launch_decode_bed(const uint8_t* bed, int64_t bytes_per_col, int n, int p, int8_t* dst, int64_t ld, int missing, int* bad,
                       unsigned long long* nmiss, cudaStream_t st) {
  decode_bed_kernel<<<(unsigned)p, 256, 0, st>>>(bed, bytes_per_col, n, dst, ld, missing, bad, nmiss);
}

__global__ void check_range_kernel(const int8_t* __restrict__ src, int64_t ld, int n, int lo, int hi, int* bad) {
  const int8_t* s = src + (int64_t)blockIdx.x * ld;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
    const int v = s[i];
    if (v < lo || v > hi) atomicExch(bad, 1);
  }
}
void launch_check_range_i8(const int8_t* src, int64_t ld, int n, int p, int lo, int hi, int* bad, cudaStream_t st) {
  dim3 grid((unsigned)p, (unsigned)((n + 1023) / 1024 > 64 ? 64 : (n + 1023) / 1024));  // columns on x (p can exceed the 65,535 limit of y)
  check_range_kernel<<<grid, 256, 0, st>>>(src, ld, n, lo, hi, bad);
}

__global__ void zero_pad_kernel(int8_t* x, int64_t ld, int n) {
  int8_t* c = x + (int64_t)blockIdx.x * ld;
  for (int i = n + threadIdx.x; i < ld; i += blockDim.x) c[i] = 0;
}
void launch_zero_pad(int8_t* x, int64_t ld, int n, int p, cudaStream_t st) {
  if (ld > n) zero_pad_kernel<<<p, 128, 0, st>>>(x, ld, n);
}

// 2-bit packing: byte k of a column holds rows 4k..4k+3, row r in bits 2*(r%4).
__global__ void pack_2bit_kernel(const int8_t* __restrict__ src, int64_t ld, int n, uint8_t* __restrict__ dst,
                                 int64_t ldb, int* bad) {
  const int8_t* s = src + (int64_t)blockIdx.x * ld;
  uint8_t* d = dst + (int64_t)blockIdx.x * ldb;
  for (int k = blockIdx.y * blockDim.x + threadIdx.x; k < ldb; k += gridDim.y * blockDim.x) {
    uint32_t byte = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int r = 4 * k + q;
      int v = 0;
      if (r < n) {
        v = s[r];
        if (v < 0 || v > 2) { atomicExch(bad, 1); v = 0; }
      }
      byte |= (uint32_t)v << (2 * q);
    }
    d[k] = (uint8_t)byte;
  }
}
void launch_pack_2bit(const int8_t* src, int64_t ld, int n, int p, uint8_t* dst, int64_t ldb, int* bad, cudaStream_t st) {
  dim3 grid((unsigned)p, (unsigned)((ldb + 255) / 256 > 64 ? 64 : (ldb + 255) / 256));  // columns on x (p can exceed the 65,535 limit of y)
  pack_2bit_kernel<<<grid, 256, 0, st>>>(src, ld, n, dst, ldb, bad);
}
// Gram shadow copy: 16 rows per 32-bit word, interleaved so that the kernel expands it with one shift and one mask per
// output word: the code of row 16g + 4q + k sits at bits [8k + 2q, 8k + 2q + 1], i.e. (w >> 2q) & 0x03030303 is the four
// bytes of rows 4q .. 4q+3.  Codes must be 0..3 (else *bad).
__global__ void pack_2bit_gram_kernel(const int8_t* __restrict__ src, int64_t ld, uint32_t* __restrict__ dst, int64_t ldw,
                                      int* bad) {
  const uint4* s = reinterpret_cast<const uint4*>(src + (int64_t)blockIdx.x * ld);
  uint32_t* d = dst + (int64_t)blockIdx.x * ldw;
  for (int64_t g = blockIdx.y * blockDim.x + threadIdx.x; g < ldw; g += (int64_t)gridDim.y * blockDim.x) {
    const uint4 v = s[g];  // rows 16g .. 16g+15 (pad rows are zero)
    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
    uint32_t out = 0, any = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      any |= w4[q] & 0xFCFCFCFCu;
      out |= (w4[q] & 0x03030303u) << (2 * q);
    }
    if (any) atomicExch(bad, 1);
    d[g] = out;
  }
}
void launch_pack_2bit_gram(const int8_t* src, int64_t ld, int p, uint8_t* dst, int* bad, cudaStream_t st) {
  const int64_t ldw = ld / 16;
  dim3 grid((unsigned)p, (unsigned)((ldw + 255) / 256 > 64 ? 64 : (ldw + 255) / 256));  // columns on x (p can exceed the 65,535 limit of y)
  pack_2bit_gram_kernel<<<grid, 256, 0, st>>>(src, ld, reinterpret_cast<uint32_t*>(dst), ldw, bad);
}
__global__ void unpack_2bit_kernel(const uint8_t* __restrict__ src, int64_t ldb, int n, int8_t* __restrict__ dst,
                                   int64_t ld) {
  const uint8_t* s = src + (int64_t)blockIdx.x * ldb;
  int8_t* d = dst + (int64_t)blockIdx.x * ld;
  for (int r = blockIdx.y * blockDim.x + threadIdx.x; r < ld; r += gridDim.y * blockDim.x)
    d[r] = (r < n) ? (int8_t)((s[r >> 2] >> (2 * (r & 3))) & 3) : (int8_t)0;
}
void launch_unpack_2bit(const uint8_t* src, int64_t ldb, int n, int p, int8_t* dst, int64_t ld, cudaStream_t st) {
  dim3 grid((unsigned)p, (unsigned)((ld + 1023) / 1024 > 64 ? 64 : (ld + 1023) / 1024));  // columns on x (p can exceed the 65,535 limit of y)
  unpack_2bit_kernel<<<grid, 256, 0, st>>>(src, ldb, n, dst, ld);
}

// Column statistics, one CTA per column, dp4a for sum and sum of squares (int32 partials are safe:
// 127^2 * 16 rows per lane-iteration, promoted to int64 before the block reduction).
template <bool MASKED>
__global__ void __launch_bounds__(256) col_stats_kernel(GenoView g, const uint8_t* __restrict__ mask, long long* xx,
                                                        long long* sx) {
  const int j = blockIdx.x;
  long long s1 = 0, s2 = 0;
  if (g.storage == 0) {
    const uint4* col = reinterpret_cast<const uint4*>(g.x8 + (int64_t)j * g.ld);
    const int nv = (int)(g.ld >> 4);
    for (int v = threadIdx.x; v < nv; v += blockDim.x) {
      uint4 w = __ldg(col + v);
      if (MASKED) {  // mask bytes are 0/1; 0x01*0xFF = 0xFF selects the genotype byte
        const uint4 m = __ldg(reinterpret_cast<const uint4*>(mask) + v);
        w.x &= m.x * 0xFFu; w.y &= m.y * 0xFFu; w.z &= m.z * 0xFFu; w.w &= m.w * 0xFFu;
      }
      int a = 0, q = 0;
      a = __dp4a((int)w.x, 0x01010101, a); a = __dp4a((int)w.y, 0x01010101, a);
      a = __dp4a((int)w.z, 0x01010101, a); a = __dp4a((int)w.w, 0x01010101, a);
      q = __dp4a((int)w.x, (int)w.x, q); q = __dp4a((int)w.y, (int)w.y, q);
      q = __dp4a((int)w.z, (int)w.z, q); q = __dp4a((int)w.w, (int)w.w, q);
      s1 += a; s2 += q;
    }
  } else {
    const uint8_t* col = g.x2 + (int64_t)j * g.ldb;
    for (int k = threadIdx.x; k < g.ldb; k += blockDim.x) {
      const uint32_t byte = col[k];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int r = 4 * k + q;
        int v = (byte >> (2 * q)) & 3;
        if (MASKED && r < g.n && !mask[r]) v = 0;
        s1 += v; s2 += v * v;
      }
    }
  }
  __shared__ long long r1[8], r2[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = s1; r2[threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long a = 0, q = 0;
    for (int w = 0; w < 8; w++) { a += r1[w]; q += r2[w]; }
    sx[j] = a; xx[j] = q;
  }
}
void launch_col_stats(const GenoView& g, long long* xx, long long* sx, cudaStream_t st) {
  col_stats_kernel<false><<<g.p, 256, 0, st>>>(g, nullptr, xx, sx);
}
void launch_col_stats_masked(const GenoView& g, const uint8_t* mask, long long* xx, long long* sx, cudaStream_t st) {
  col_stats_kernel<true><<<g.p, 256, 0, st>>>(g, mask, xx, sx);
}

// float32 store: xx_j = sum x^2, sx_j = sum x in double, fixed order (one CTA per marker)
__global__ void __launch_bounds__(256) col_stats_f32_kernel(GenoView g, const uint8_t* __restrict__ mask, double* __restrict__ xx,
                                                            double* __restrict__ sx) {
  __shared__ double sh1[8], sh2[8];
  const int j = blockIdx.x, tid = threadIdx.x;
  const float* col = g.xf + (int64_t)j * g.ld;
  double s1 = 0, s2 = 0;
  for (int i = tid; i < g.n; i += 256) {
    if (mask && !mask[i]) continue;
    const double v = (double)col[i];
    s1 += v; s2 += v * v;
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  if ((tid & 31) == 0) { sh1[tid >> 5] = s1; sh2[tid >> 5] = s2; }
  __syncthreads();
  if (tid == 0) {
    double a = 0, c = 0;
    for (int w = 0; w < 8; w++) { a += sh1[w]; c += sh2[w]; }
    sx[j] = a; xx[j] = c;
  }
}
void launch_col_stats_f32(const GenoView& g, const uint8_t* mask, double* xx, double* sx, cudaStream_t st) {
  col_stats_f32_kernel<<<g.p, 256, 0, st>>>(g, mask, xx, sx);
}
// Row multiplicities (KMUP2 on rows sampled with replacement, R/wgr.R:68 with rp = TRUE): xx_j = sum_i c_i x_ij^2, sx_j = sum_i c_i x_ij
// for any store, in double (exact for the integer stores: every partial sum is an integer below 2^53).  One CTA per marker.
__global__ void __launch_bounds__(256) col_stats_cnt_kernel(GenoView g, const uint8_t* __restrict__ cnt, double* __restrict__ xx,
                                                            double* __restrict__ sx) {
  __shared__ double sh1[8], sh2[8];
  const int j = blockIdx.x, tid = threadIdx.x;
  double s1 = 0, s2 = 0;
  for (int i = tid; i < g.n; i += 256) {
    const double c = (double)cnt[i];
    if (c == 0.0) continue;
    double v;
    if (g.storage == 0) v = (double)g.x8[(int64_t)j * g.ld + i];
    else if (g.storage == 1) v = (double)((g.x2[(int64_t)j * g.ldb + (i >> 2)] >> (2 * (i & 3))) & 3);
    else v = (double)g.xf[(int64_t)j * g.ld + i];
    s1 += c * v; s2 += c * v * v;
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  if ((tid & 31) == 0) { sh1[tid >> 5] = s1; sh2[tid >> 5] = s2; }
  __syncthreads();
  if (tid == 0) {
    double a = 0, c = 0;
    for (int w = 0; w < 8; w++) { a += sh1[w]; c += sh2[w]; }
    sx[j] = a; xx[j] = c;
  }
}
void launch_col_stats_cnt(const GenoView& g, const uint8_t* cnt, double* xx, double* sx, cudaStream_t st) {
  col_stats_cnt_kernel<<<g.p, 256, 0, st>>>(g, cnt, xx, sx);
}
__global__ void __launch_bounds__(256) d_to_float_kernel(const double* __restrict__ src, float* __restrict__ dst, int n) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) dst[i] = (float)src[i];
}
void launch_d_to_float(const double* src, float* dst, int n, cudaStream_t st) { d_to_float_kernel<<<(n + 255) / 256, 256, 0, st>>>(src, dst, n); }

// hat = mu + X b.  CTA (x = row tile of 256*16 rows, y = column split): partial[y][row] = sum over its
// columns; then a second kernel adds the splits in a fixed order (deterministic).
__global__ void __launch_bounds__(256) gemv_partial_kernel(GenoView g, const float* __restrict__ b, float* __restrict__ work,
                                                           int splits) {
  const int row0 = (blockIdx.x * 256 + threadIdx.x) * 16;
  if (row0 >= g.ld) return;
  float acc[16];
#pragma unroll
  for (int q = 0; q < 16; q++) acc[q] = 0.0f;
  for (int j = blockIdx.y; j < g.p; j += splits) {
    const float bj = __ldg(b + j);
    if (bj == 0.0f) continue;
    if (g.storage == 2) {  // float32 store
      const float4* v = reinterpret_cast<const float4*>(g.xf + (int64_t)j * g.ld + row0);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const float4 x = __ldg(v + q);
        acc[4 * q + 0] = fmaf(x.x, bj, acc[4 * q + 0]); acc[4 * q + 1] = fmaf(x.y, bj, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(x.z, bj, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(x.w, bj, acc[4 * q + 3]);
      }
      continue;
    }
    uint32_t w[4];
    if (g.storage == 0) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(g.x8 + (int64_t)j * g.ld + row0));
      w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    } else {
      const uint32_t pk = __ldg(reinterpret_cast<const uint32_t*>(g.x2 + (int64_t)j * g.ldb + (row0 >> 2)));
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const uint32_t byte = (pk >> (8 * q)) & 0xFFu;
        w[q] = (byte & 3u) | (((byte >> 2) & 3u) << 8) | (((byte >> 4) & 3u) << 16) | (((byte >> 6) & 3u) << 24);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const uint32_t x = w[q] ^ 0x80808080u;
#pragma unroll
      for (int t = 0; t < 4; t++) acc[4 * q + t] = fmaf(byte_to_float(x, t), bj, acc[4 * q + t]);
    }
  }
  float* out = work + (int64_t)blockIdx.y * g.ld + row0;
#pragma unroll
  for (int q = 0; q < 16; q += 4) *reinterpret_cast<float4*>(out + q) = make_float4(acc[q], acc[q + 1], acc[q + 2], acc[q + 3]);
}
__global__ void gemv_reduce_kernel(const float* __restrict__ work, int64_t ld, int n, int splits, const float* mu,
                                   float* __restrict__ hat) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.0f;
  for (int k = 0; k < splits; k++) s += work[(int64_t)k * ld + i];
  hat[i] = s + mu[0];
}
void launch_gemv_hat(const GenoView& g, const float* b, const float* mu_dev, float* hat, float* work, int splits,
                     cudaStream_t st) {
  dim3 grid((unsigned)((g.ld + 4095) / 4096), splits);
  gemv_partial_kernel<<<grid, 256, 0, st>>>(g, b, work, splits);
  gemv_reduce_kernel<<<(g.n + 255) / 256, 256, 0, st>>>(work, g.ld, g.n, splits, mu_dev, hat);
}

}  // namespace bwgr
