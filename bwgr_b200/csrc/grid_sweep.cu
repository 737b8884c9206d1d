// grid_sweep.cu -- the third kernel family ("grid"): the reference's per-marker step (dot -> rule -> residual update), literally, with the
// individuals spread over the whole GPU.  It takes what the other two families cannot: any n on one GPU (the small-n family needs the
// residual of a system in ONE SM's shared memory, the blocked family at most 512 rows per worker, i.e. n <= ~75k), row-masked systems
// (CV folds) at any n, every rule of common.cuh.  Up to 32 systems share one pass over the genotypes.
//
// One persistent cooperative grid; CTA c keeps its row slab of the residuals (and masks) of all systems in shared memory for the whole
// sweep and streams its slab of the genotype columns in marker order through a cp.async ring.  Per marker: slab dot products -> ONE
// grid-wide sum through L2 -> the rule, evaluated by every CTA alike -> rank-one update of the slab.  The grid sum: each CTA adds its
// fixed-point partial to the marker's own 64-bit accumulator word with one atomic per system; the low byte of a word counts the
// arrivals, so the data is its own flag (one L2 hop, no fence, no barrier).  Integer sums are order-free: every CTA reads bit-identical
// totals, takes bit-identical decisions, and a fit is bit-reproducible.  Same-address atomics serialise in L2, so CTA c adds to copy
// c % 8 of the word.  The per-marker chain is latency-bound (~2-3 us per marker): this family is the fallback, not the fast path.
// HBM traffic per sweep: n p bytes.
#include "kernels.h"

namespace bwgr {

namespace {

constexpr int kT = 256;  // threads per CTA
constexpr int kD = 4;    // markers in flight
constexpr int kC = kGridCopies;

__device__ __forceinline__ void gcp16(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void gcp4(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void gcp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void gcp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ unsigned long long gld_relaxed(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

struct GridSmem {
  float* E;          // [ns][rp]
  uint8_t* M;        // [ns][rp] (masked fits)
  unsigned char* xs[kD];  // [rp] int8 or [rp] float
  float* vin[kD];    // b0[32] | vbj[32] | xx[32] | xx2
  SysScalars* sc;    // [ns]
  MarkerDraws* dr;   // [2][32]
  int* Js;           // [kD]
};

__host__ __device__ inline size_t grid_carve(unsigned char* base, int ns, int rp, bool masked, int xbytes, GridSmem* s) {
  size_t o = 0;
  auto take = [&](size_t bytes) { unsigned char* q = base ? base + o : nullptr; o = (o + bytes + 15) & ~(size_t)15; return q; };
  float* E = reinterpret_cast<float*>(take(sizeof(float) * (size_t)ns * rp));
  unsigned char* xs[kD];
  for (int d = 0; d < kD; d++) xs[d] = take((size_t)rp * xbytes);
  uint8_t* M = reinterpret_cast<uint8_t*>(take(masked ? (size_t)ns * rp : 0));
  float* vin[kD];
  for (int d = 0; d < kD; d++) vin[d] = reinterpret_cast<float*>(take(sizeof(float) * 100));
  SysScalars* sc = reinterpret_cast<SysScalars*>(take(sizeof(SysScalars) * (size_t)ns));
  MarkerDraws* dr = reinterpret_cast<MarkerDraws*>(take(sizeof(MarkerDraws) * 64));
  int* Js = reinterpret_cast<int*>(take(sizeof(int) * kD));
  if (s) {
    s->E = E; s->M = M; s->sc = sc; s->dr = dr; s->Js = Js;
    for (int d = 0; d < kD; d++) { s->xs[d] = xs[d]; s->vin[d] = vin[d]; }
  }
  return o;
}

// XT = int8_t (integer store) or float (real-valued store: NA-imputed / centred genotypes, what the reference holds as MatrixXf)
template <int MODEL, class XT>
__global__ void __launch_bounds__(kT, 1) grid_sweep_kernel(GridArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_abort;
  __shared__ float s_de[32];
  __shared__ float s_wred[8][33];
  GridSmem s;
  const int ns = a.nsys, rp = a.rows_per_cta, p = a.g.p;
  const bool masked = a.mask != nullptr;
  grid_carve(smem_raw, ns, rp, masked, (int)sizeof(XT), &s);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, G = (int)gridDim.x, cta = blockIdx.x;
  const int64_t ld = a.g.ld, r0 = (int64_t)cta * rp;
  for (int q = tid; q < ns * rp; q += kT) {
    const int t = q / rp, i = q - t * rp;
    const bool in = r0 + i < ld;
    s.E[q] = in ? a.e[(size_t)t * ld + r0 + i] : 0.0f;
    if (masked) s.M[q] = in ? a.mask[(size_t)t * ld + r0 + i] : (uint8_t)0;
  }
  for (int q = tid; q < kD * rp * (int)sizeof(XT); q += kT) s.xs[0][q] = 0;  // the ring slots are contiguous (rp is a multiple of 16)
  if (tid < ns) s.sc[tid] = a.sc[tid];
  if (tid == 0) s_abort = 0;
  const int nch = rp * (int)sizeof(XT) / 16;  // 16-byte chunks of a slab
  constexpr int kRowsPerChunk = 16 / (int)sizeof(XT);
  const float* xxbase = a.xx;
  int Jnext = a.perm ? a.perm[0] : 0;
  auto prefetch = [&](int m) {
    if (m < p) {
      const int J = Jnext, slot = m % kD;
      if (m + 1 < p) Jnext = a.perm ? a.perm[m + 1] : m + 1;
      const unsigned char* col = sizeof(XT) == 1 ? reinterpret_cast<const unsigned char*>(a.g.x8 + (int64_t)J * ld + r0)
                                                 : reinterpret_cast<const unsigned char*>(a.g.xf + (int64_t)J * ld + r0);
      for (int c = tid; c < nch; c += kT)
        if (r0 + (int64_t)kRowsPerChunk * c < ld) gcp16(s.xs[slot] + 16 * c, col + 16 * c);
      float* v = s.vin[slot];
      if (tid < 32) { if (tid < ns) gcp4(v + tid, a.b + (size_t)tid * p + J); }
      else if (tid < 64) { const int t = tid - 32; if (t < ns && a.vbv) gcp4(v + 32 + t, a.vbv + (size_t)t * p + J); }
      else if (tid < 96) { const int t = tid - 64; if (t < ns) gcp4(v + 64 + t, xxbase + (a.xx_per_sys ? (size_t)t * p : 0) + J); }
      else if (tid == 96 && a.xx2) gcp4(v + 96, a.xx2 + J);
      if (tid == 0) s.Js[slot] = J;
    }
    gcp_commit();
  };
  __syncthreads();  // the zeroed ring slots are in place before any cp.async lands in them
  for (int m = 0; m < kD - 1; m++) prefetch(m);
  __syncthreads();
  if (model_is_gibbs(MODEL) && warp == 1 && lane < ns)  // draws of marker 0
    s.dr[lane] = marker_draws(MODEL, (uint32_t)s.Js[0], (uint32_t)s.sc[lane].sweep, (uint32_t)(a.chain0 + lane), s.sc[lane].df, a.seed_lo, a.seed_hi);
  const unsigned long long t_start = gtimer();
  const double inv_q = (double)a.g_quantum;  // value of one fixed-point unit
  const double qinv = 1.0 / (double)a.g_quantum;
  for (int m = 0; m < p; m++) {
    gcp_wait<kD - 2>();
    __syncthreads();  // marker m's slot has landed; step m - 1 is finished by every thread
    prefetch(m + kD - 1);
    const int slot = m % kD;
    const XT* xs = reinterpret_cast<const XT*>(s.xs[slot]);
    const float* vin = s.vin[slot];
    // ---- slab dot products g_s = x'e_s
    if (ns == 1) {
      float acc = 0.0f;
      if (masked) for (int i = tid; i < rp; i += kT) acc = fmaf((float)xs[i] * (float)s.M[i], s.E[i], acc);
      else for (int i = tid; i < rp; i += kT) acc = fmaf((float)xs[i], s.E[i], acc);
      acc = warp_sum(acc);
      if (lane == 0) s_wred[warp][0] = acc;
    } else {
      for (int t0 = 0; t0 < ns; t0 += 8) {
        float acc[8];
#pragma unroll
        for (int q = 0; q < 8; q++) acc[q] = 0.0f;
        for (int i = tid; i < rp; i += kT) {
          const float x = (float)xs[i];
#pragma unroll
          for (int q = 0; q < 8; q++)
            if (t0 + q < ns) acc[q] = fmaf(masked ? x * (float)s.M[(t0 + q) * rp + i] : x, s.E[(t0 + q) * rp + i], acc[q]);
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
          if (t0 + q < ns) {
            const float v = warp_sum(acc[q]);
            if (lane == 0) s_wred[warp][t0 + q] = v;
          }
        }
      }
    }
    __syncthreads();
    if (warp == 0) {
      const bool on = lane < ns;
      if (on) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < 8; w++) v += (double)s_wred[w][lane];
        v *= qinv;
        if (!(fabs(v) < 4503599627370496.0)) atomicExch(a.err, 4);  // 2^52: the fixed-point range of g is exceeded
        atomicAdd(a.acc + ((size_t)m * kC + (cta & (kC - 1))) * 32 + lane, ((unsigned long long)__double2ll_rn(v) << 8) + 1ull);
      }
      // the marker's totals: a word is complete when its low byte counts all the CTAs that add to that copy
      const unsigned long long* accw = a.acc + (size_t)m * kC * 32 + (on ? lane : 0);
      unsigned long long w[kC];
      unsigned int spins = 0;
      for (;;) {
        bool pending = false;
#pragma unroll
        for (int c = 0; c < kC; c++) w[c] = gld_relaxed(accw + c * 32);
#pragma unroll
        for (int c = 0; c < kC; c++) pending |= (int)(w[c] & 0xffull) != (G + kC - 1 - c) / kC;
        if (!__any_sync(0xffffffffu, on && pending)) break;
        if ((++spins & 0xfffu) == 0) {
          int ab = 0;
          if (lane == 0) {
            if (*reinterpret_cast<volatile int*>(a.err) != 0) ab = 1;
            else if (gtimer() - t_start > 120000000000ull) { atomicExch(a.err, 3); ab = 1; }  // two minutes without the grid
            if (ab) s_abort = 1;
          }
          if (__shfl_sync(0xffffffffu, ab, 0)) break;
        }
      }
      if (on) {
        long long tot = 0;
#pragma unroll
        for (int c = 0; c < kC; c++) tot += (long long)w[c] >> 8;
        const float g = (float)((double)tot * inv_q);
        const SysScalars& sc = s.sc[lane];
        float de = 0.0f;
        if (!sc.done) {
          MarkerDraws dr;
          if (model_is_gibbs(MODEL)) dr = s.dr[(m & 1) * 32 + lane];
          else { dr.z1 = dr.z2 = dr.u = 0.0f; dr.chi = 1.0f; }
          const float vbj = a.vbv ? vin[32 + lane] : 0.0f;
          const RuleOut r = marker_rule<MODEL>(g, vin[64 + lane], vin[lane], vbj, sc, dr, a.xx2 ? vin[96] : 0.0f);
          de = r.de;
          if (cta == 0) {
            const size_t q = (size_t)lane * p + s.Js[slot];
            a.b[q] = r.b;
            if (model_has_d(MODEL) && a.d) a.d[q] = r.d;
            if (model_rule_writes_vbj(MODEL) && a.vbv) a.vbv[q] = r.vbj;
          }
        }
        s_de[lane] = de;
      }
    } else if (model_is_gibbs(MODEL) && warp == 1 && lane < ns && m + 1 < p) {  // the next marker's draws, off the chain
      s.dr[((m + 1) & 1) * 32 + lane] = marker_draws(MODEL, (uint32_t)s.Js[(m + 1) % kD], (uint32_t)s.sc[lane].sweep, (uint32_t)(a.chain0 + lane),
                                                     s.sc[lane].df, a.seed_lo, a.seed_hi);
    }
    __syncthreads();
    if (s_abort) return;  // the host reports the error flag; the residuals of this launch are not written back
    // ---- e_s -= x de_s (on the rows the system uses)
    for (int t = 0; t < ns; t++) {
      const float de = s_de[t];
      if (de == 0.0f) continue;
      float* Et = s.E + (size_t)t * rp;
      if (masked) {
        const uint8_t* Mt = s.M + (size_t)t * rp;
        for (int i = tid; i < rp; i += kT) Et[i] = fmaf(-(float)xs[i] * (float)Mt[i], de, Et[i]);
      } else {
        for (int i = tid; i < rp; i += kT) Et[i] = fmaf(-(float)xs[i], de, Et[i]);
      }
    }
  }
  gcp_wait<0>();
  __syncthreads();
  for (int q = tid; q < ns * rp; q += kT) {
    const int t = q / rp, i = q - t * rp;
    if (r0 + i < ld) a.e[(size_t)t * ld + r0 + i] = s.E[q];
  }
}

template <int MODEL>
cudaError_t launch_grid_model(const GridArgs& a, int grid, cudaStream_t st) {
  const bool real = a.g.storage == 2;
  const size_t smem = grid_sweep_smem(a.nsys, a.rows_per_cta, a.mask != nullptr, real);
  const void* fn = real ? reinterpret_cast<const void*>(grid_sweep_kernel<MODEL, float>) : reinterpret_cast<const void*>(grid_sweep_kernel<MODEL, int8_t>);
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  GridArgs args = a;
  void* params[] = {&args};
  return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kT), params, smem, st);
}

}  // namespace

size_t grid_sweep_smem(int nsys, int rows_per_cta, bool masked, bool real_store) {
  return grid_carve(nullptr, nsys, rows_per_cta, masked, real_store ? 4 : 1, nullptr) + 16;
}

cudaError_t launch_grid_sweep(const GridArgs& a, int grid, cudaStream_t st) {
  switch (rule_model(a.model)) {
    case M_EMRR: return launch_grid_model<M_EMRR>(a, grid, st);
    case M_EMBA: return launch_grid_model<M_EMBA>(a, grid, st);
    case M_EMBB: return launch_grid_model<M_EMBB>(a, grid, st);
    case M_EMBC: return launch_grid_model<M_EMBC>(a, grid, st);
    case M_EMBL: return launch_grid_model<M_EMBL>(a, grid, st);
    case M_EMEN: return launch_grid_model<M_EMEN>(a, grid, st);
    case M_EMDE: return launch_grid_model<M_EMDE>(a, grid, st);
    case M_LASSO: return launch_grid_model<M_LASSO>(a, grid, st);
    case M_BL: return launch_grid_model<M_BL>(a, grid, st);
    case M_BDPI: return launch_grid_model<M_BDPI>(a, grid, st);
    case M_BRR: return launch_grid_model<M_BRR>(a, grid, st);
    case M_BA: return launch_grid_model<M_BA>(a, grid, st);
    case M_BB: return launch_grid_model<M_BB>(a, grid, st);
    case M_BC: return launch_grid_model<M_BC>(a, grid, st);
    case M_KMUP: return launch_grid_model<M_KMUP>(a, grid, st);
    case M_KMUP2: return launch_grid_model<M_KMUP2>(a, grid, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace bwgr
