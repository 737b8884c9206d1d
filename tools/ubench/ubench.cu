// ubench.cu -- B200 micro-benchmarks behind the design of the pipelined sweep (sweep_pipe.cu).  Stand-alone binary:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench/ubench tools/ubench/ubench.cu
// A: issue / completion time of the tiny-N tcgen05.mma groups of a worker (dependent vs independent accumulators)
// B: grid-wide flag exchange through L2 (publish -> all workers -> reduction back to one CTA), several protocols
// C: thread-block-cluster co-residency with ~200 KB of shared memory per CTA and DSMEM st.async round trips
// Every wait is bounded: a protocol bug ends the kernel with a flag instead of hanging the GPU.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr uint32_t kSpin = 1u << 22;
constexpr int kAtomBytes = 128 * 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t s = 0; s < kSpin; s++) if (mbar_try(bar, parity)) return true;
  return false;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(kAtomBytes >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t idesc_i8(int N, int a_mn) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}
__host__ __device__ constexpr uint32_t idesc_f8(int N, int a_mn) {  // E4M3 x E4M3 -> F32
  return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}
template <int F8>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if (F8)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void ld_relaxed_v2(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

// =====================================================================================================================
// A: tcgen05 issue patterns.  pattern: 0 = G-like, 12 K-major MMAs into ONE accumulator; 1 = one accumulator per atom (3 chains
// of 4, atom-major order); 2 = 12 accumulators; 3 = U-like (A MN-major), 3 accumulators, atom-major order (4 dependent in a
// row); 4 = U-like, k4-major order (chains interleaved); 5 = U-like, 12 accumulators; 6 = G-like 3 chains interleaved
// =====================================================================================================================
template <int F8>
__global__ void __launch_bounds__(128, 1) mma_issue_kernel(int N, int pattern, int reps, long long* out) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* Xs = base;                       // 3 atoms
  unsigned char* EL = base + 3 * kAtomBytes;      // B operand, up to N = 128: 3 atoms x (N/8) KB
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (3 * kAtomBytes + 3 * 16 * 1024) / 16; i += 128) reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (warp == 0) {
    const uint32_t idg = F8 ? idesc_f8(N, 0) : idesc_i8(N, 0), idu = F8 ? idesc_f8(N, 1) : idesc_i8(N, 1);
    const int ncol = N;  // TMEM columns per accumulator
    for (int r = 0; r < reps; r++) {
      __syncwarp();
      long long t0 = 0, t1 = 0, t2 = 0;
      const bool el = elect_one();
      if (el) {
        t0 = clock64();
        if (pattern <= 2 || pattern == 6) {
          for (int i = 0; i < 12; i++) {
            const int at = pattern == 6 ? i % 3 : i / 4, k4 = pattern == 6 ? i / 3 : i % 4;
            const uint64_t ad = desc_k_sw128(smem_u32(Xs + (size_t)at * kAtomBytes)) + (uint64_t)(2 * k4);
            const uint64_t bd = desc_k_sw128(smem_u32(EL + (size_t)at * (N / 8) * 1024)) + (uint64_t)(2 * k4);
            const int accum = pattern == 0 ? 0 : (pattern == 1 || pattern == 6) ? at : (at * 4 + k4);
            const bool first = pattern == 0 ? i == 0 : (pattern == 1 || pattern == 6) ? k4 == 0 : true;
            if ((accum + 1) * ncol <= 512) umma<F8>(tmem + (uint32_t)(accum * ncol), ad, bd, idg, first ? 0u : 1u);
          }
        } else {
          for (int i = 0; i < 12; i++) {
            const int at = pattern == 4 ? i % 3 : i / 4, k4 = pattern == 4 ? i / 3 : i % 4;
            const uint64_t ad = desc_mn_sw128(smem_u32(Xs + (size_t)at * kAtomBytes)) + (uint64_t)(k4 * (4096 >> 4));
            const uint64_t bd = desc_k_sw128(smem_u32(EL)) + (uint64_t)(2 * k4);
            const int accum = pattern == 5 ? at * 4 + k4 : at;
            const bool first = pattern == 5 ? true : k4 == 0;
            if ((accum + 1) * ncol <= 512) umma<F8>(tmem + (uint32_t)(accum * ncol), ad, bd, idu, first ? 0u : 1u);
          }
        }
        t1 = clock64();
        umma_commit(&bar);
      }
      __syncwarp();
      const bool ok = mbar_wait(&bar, (uint32_t)r & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (el) {
        t2 = clock64();
        out[2 * r] = t1 - t0;
        out[2 * r + 1] = ok ? t2 - t0 : -1;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

static void run_A() {
  long long* d;
  const int reps = 64;
  CK(cudaMalloc(&d, sizeof(long long) * 2 * reps));
  const size_t smem = 3 * kAtomBytes + 3 * 16 * 1024 + 1024;
  CK(cudaFuncSetAttribute(mma_issue_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(mma_issue_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  printf("A: 12 tcgen05.mma (M=128, K=32) -- cycles to issue / cycles until the commit arrives (median over %d reps)\n", reps);
  for (int f8 = 0; f8 < 2; f8++)
    for (int N : {16, 32, 64, 128})
      for (int pat = 0; pat <= 6; pat++) {
        if ((pat == 2 || pat == 5) && 12 * N > 512 && N > 32) continue;
        if (f8) mma_issue_kernel<1><<<1, 128, smem>>>(N, pat, reps, d);
        else mma_issue_kernel<0><<<1, 128, smem>>>(N, pat, reps, d);
        CK(cudaDeviceSynchronize());
        std::vector<long long> h(2 * reps);
        CK(cudaMemcpy(h.data(), d, sizeof(long long) * 2 * reps, cudaMemcpyDeviceToHost));
        std::vector<long long> a, b;
        for (int r = 8; r < reps; r++) { a.push_back(h[2 * r]); b.push_back(h[2 * r + 1]); }
        std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end());
        printf("  kind=%s N=%3d pattern=%d  issue %5lld  done %5lld   (min done %lld)\n", f8 ? "f8" : "i8", N, pat, a[a.size() / 2], b[b.size() / 2], b[0]);
      }
  cudaFree(d);
}

// =====================================================================================================================
// B: grid-wide exchange through L2.  CTA 0 = solver, CTAs 1..G-1 = workers.  One iteration = publish 129 words -> every worker
// sees them -> every worker answers -> the solver has all answers.  mode:
//   0  answer = one word per worker in a contiguous array, solver warp polls it (pure 2-hop flag latency)
//   1  answer = 128 words per worker (1 KB rows), ALL read directly by the solver CTA's 1024 threads (one hop)
//   2  answer = 128 words per worker, two-hop tree as in sweep_pipe v5 (reducer warp in worker m % W sums marker m)
//   3  answer = red.add on one counter
//   4  as 1 but only `nread` of the 128 words per worker are read by this CTA (emulates a solver cluster of 128/nread CTAs)
// poll_mode (workers): 0 = 4 words per lane + scale word (v5), 1 = lane 0 polls one flag word written last after a fence
// =====================================================================================================================
__global__ void __launch_bounds__(1024, 1) flag_kernel(int mode, int poll_mode, int nread, int iters, unsigned long long* pub, unsigned long long* ans,
                                                        unsigned long long* red, unsigned int* counter, long long* out, int* fail) {
  extern __shared__ unsigned char smem_raw[];
  (void)smem_raw;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, W = G - 1;
  __shared__ int s_ok;
  __shared__ long long s_sum[32];
  bool dead = false;
  if (blockIdx.x == 0) {
    long long t0 = 0;
    for (int it = 0; it < iters; it++) {
      const unsigned long long tag = (unsigned long long)it + 1;
      if (it == 8 && tid == 0) t0 = clock64();
      // publish
      if (tid < 128) st_relaxed_u64(pub + tid, ((unsigned long long)tid << 32) | tag);
      if (poll_mode == 1) {
        __syncthreads();
        if (tid == 0) { __threadfence(); st_relaxed_u64(pub + 128, tag); }
      } else if (tid == 0) st_relaxed_u64(pub + 128, tag);
      // collect
      if (mode == 0) {
        if (warp == 0) {
          uint32_t spin = 0;
          while (!dead) {
            bool ok = true;
            for (int k = 0; k < 5; k++) { const int w = 32 * k + lane; if (w < W) ok = ok && (ld_relaxed_u64(ans + w) == tag); }
            if (__all_sync(0xffffffffu, ok)) break;
            if (++spin > kSpin) { dead = true; *fail = 1; }
          }
        }
        __syncthreads();
      } else if (mode == 1 || mode == 4) {
        const int per = mode == 4 ? nread : 128;          // words per worker read here
        const int total2 = W * per / 2;                    // 16-byte loads
        long long sum = 0;
        for (int i = tid; i < total2; i += 1024) {
          const int w = (2 * i) / per, m = (2 * i) % per;
          const unsigned long long* p = ans + (size_t)w * 128 + m;
          unsigned long long a = 0, b = 0;
          uint32_t spin = 0;
          while (!dead) {
            ld_relaxed_v2(p, a, b);
            if ((a & 0xFFFull) == (tag & 0xFFFull) && (b & 0xFFFull) == (tag & 0xFFFull)) break;
            if (++spin > kSpin) { dead = true; *fail = 2; }
          }
          sum += (long long)(a >> 12) + (long long)(b >> 12);
        }
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) s_sum[warp] = sum;
        __syncthreads();
      } else if (mode == 2) {
        if (tid < 128) {
          uint32_t spin = 0;
          while (!dead) {
            const unsigned long long v = ld_relaxed_u64(red + tid);
            if (__all_sync(0xffffffffu, (v & 0xFFFull) == (tag & 0xFFFull))) break;
            if (++spin > kSpin) { dead = true; *fail = 3; }
          }
        }
        __syncthreads();
      } else {
        if (tid == 0) {
          uint32_t spin = 0;
          while (*reinterpret_cast<volatile unsigned int*>(counter) < (unsigned int)W * (unsigned int)(it + 1)) if (++spin > kSpin) { *fail = 4; break; }
        }
        __syncthreads();
      }
    }
    if (tid == 0) out[0] = clock64() - t0;
  } else {
    const int w = blockIdx.x - 1;
    for (int it = 0; it < iters; it++) {
      const unsigned long long tag = (unsigned long long)it + 1;
      if (warp == 0) {
        uint32_t spin = 0;
        if (poll_mode == 0) {
          while (!dead) {
            bool ok = true;
            for (int t = 0; t < 4; t++) ok = ok && ((ld_relaxed_u64(pub + 32 * t + lane) & 0xFFFFFFFFull) == tag);
            if (lane == 0) ok = ok && (ld_relaxed_u64(pub + 128) == tag);
            if (__all_sync(0xffffffffu, ok)) break;
            if (++spin > kSpin) { dead = true; *fail = 5; }
          }
        } else {
          if (lane == 0) while (ld_relaxed_u64(pub + 128) != tag) if (++spin > kSpin) { dead = true; *fail = 6; break; }
          __syncwarp();
          __threadfence();
          unsigned long long acc = 0;
          for (int t = 0; t < 4; t++) acc += ld_relaxed_u64(pub + 32 * t + lane);
          if (acc == 0xdeadbeefdeadbeefull) *fail = 99;
        }
        if (mode == 0 && lane == 0) st_relaxed_u64(ans + w, tag);
        if (mode == 3 && lane == 0) atomicAdd(counter, 1u);
        s_ok = it;
      }
      if (mode == 1 || mode == 2 || mode == 4) {
        // the four "epilogue warps" store 128 partial words once warp 0 has the step
        __syncthreads();
        if (tid < 128) st_relaxed_u64(ans + (size_t)w * 128 + tid, ((unsigned long long)(w + tid) << 12) | (tag & 0xFFFull));
        if (mode == 2 && warp == 9) {
          for (int task = w; task < 128; task += W) {
            const unsigned long long* row = ans + task;
            long long sum = 0;
            uint32_t spin = 0;
            while (!dead) {
              bool ok = true;
              sum = 0;
              for (int k = 0; k < 5; k++) {
                const int ww = 32 * k + lane;
                if (ww < W) { const unsigned long long v = ld_relaxed_u64(row + (size_t)ww * 128); ok = ok && ((v & 0xFFFull) == (tag & 0xFFFull)); sum += (long long)(v >> 12); }
              }
              if (__all_sync(0xffffffffu, ok)) break;
              if (++spin > kSpin) { dead = true; *fail = 7; }
            }
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (lane == 0) st_relaxed_u64(red + task, ((unsigned long long)sum << 12) | (tag & 0xFFFull));
          }
        }
      }
    }
  }
}

static void run_B(int sms) {
  unsigned long long *pub, *ans, *red;
  unsigned int* counter;
  long long* out;
  int* fail;
  CK(cudaMalloc(&pub, 8 * 256)); CK(cudaMalloc(&ans, 8 * 128 * 160)); CK(cudaMalloc(&red, 8 * 128)); CK(cudaMalloc(&counter, 4));
  CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&fail, 4));
  const size_t smem = 120 * 1024;  // one CTA per SM
  CK(cudaFuncSetAttribute(flag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int iters = 408;
  printf("B: publish -> all workers -> answer -> solver, grid %d, cycles per iteration\n", sms);
  struct Cfg { int mode, poll, nread; const char* name; };
  const Cfg cfgs[] = {
      {0, 0, 0, "flag word per worker, v5 poll (4 words/lane)"}, {0, 1, 0, "flag word per worker, single-flag poll"},
      {3, 0, 0, "red.add counter"},
      {1, 0, 128, "one hop: solver CTA reads all W x 128 words"}, {4, 0, 64, "one hop: W x 64 words (cluster of 2)"},
      {4, 0, 32, "one hop: W x 32 words (cluster of 4)"}, {4, 0, 16, "one hop: W x 16 words (cluster of 8)"},
      {2, 0, 0, "two-hop tree (v5)"}};
  for (const Cfg& c : cfgs) {
    CK(cudaMemset(pub, 0, 8 * 256)); CK(cudaMemset(ans, 0, 8 * 128 * 160)); CK(cudaMemset(red, 0, 8 * 128)); CK(cudaMemset(counter, 0, 4)); CK(cudaMemset(fail, 0, 4));
    int mode = c.mode, poll = c.poll, nread = c.nread, it = iters;
    void* args[] = {&mode, &poll, &nread, &it, &pub, &ans, &red, &counter, &out, &fail};
    CK(cudaLaunchCooperativeKernel((void*)flag_kernel, dim3(sms), dim3(1024), args, smem, 0));
    CK(cudaDeviceSynchronize());
    long long cyc; int f;
    CK(cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&f, fail, 4, cudaMemcpyDeviceToHost));
    printf("  %-52s %7.0f cycles/iter  fail=%d\n", c.name, (double)cyc / (iters - 8), f);
  }
}

// =====================================================================================================================
// C: clusters.  Occupancy of clusters with big CTAs, and the DSMEM round trip rank 0 -> all peers (128 words each) -> rank 0.
// =====================================================================================================================
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_u64(uint32_t raddr, unsigned long long v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(raddr), "l"(v), "r"(rbar) : "memory");
}
__global__ void __launch_bounds__(512, 1) cluster_kernel(int iters, long long* out, int* fail) {
  extern __shared__ unsigned char smem_raw[];
  cg::cluster_group cl = cg::this_cluster();
  const uint32_t rank = cl.block_rank(), csize = cl.num_blocks();
  unsigned long long* inbox = reinterpret_cast<unsigned long long*>(smem_raw);          // [16][128] (rank 0: answers; peers: slot 0)
  __shared__ uint64_t bar;
  const int tid = threadIdx.x;
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  cl.sync();
  long long t0 = 0;
  bool dead = false;
  for (int it = 0; it < iters && !dead; it++) {
    if (it == 8 && tid == 0) t0 = clock64();
    if (rank == 0) {
      if (tid == 0) mbar_expect_tx(&bar, (csize - 1) * 1024u);
      // step to every peer
      for (int i = tid; i < (int)(csize - 1) * 128; i += 512) {
        const uint32_t peer = 1 + i / 128, m = i % 128;
        st_async_u64(mapa(smem_u32(inbox + m), peer), (unsigned long long)it * 1000 + m, mapa(smem_u32(&bar), peer));
      }
      if (tid == 0 && !mbar_wait(&bar, (uint32_t)it & 1u)) { *fail = 1; }
      __syncthreads();
      if (*reinterpret_cast<volatile int*>(fail)) dead = true;
    } else {
      if (tid == 0) {
        mbar_expect_tx(&bar, 1024u);
        if (!mbar_wait(&bar, (uint32_t)it & 1u)) *fail = 2;
      }
      __syncthreads();
      if (*reinterpret_cast<volatile int*>(fail)) dead = true;
      if (tid < 128) st_async_u64(mapa(smem_u32(inbox + rank * 128 + tid), 0), inbox[tid] + rank, mapa(smem_u32(&bar), 0));
    }
  }
  if (rank == 0 && tid == 0 && blockIdx.x == 0) out[0] = clock64() - t0;
  cl.sync();
}

static void run_C() {
  printf("C: clusters\n");
  const size_t smem = 200 * 1024;
  CK(cudaFuncSetAttribute(cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  long long* out; int* fail;
  CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&fail, 4));
  for (int cs : {2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int ncl = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, cluster_kernel, &cfg);
    printf("  cluster size %2d, 512 threads, 200 KB smem: max active clusters %d (%d CTAs)  [%s]\n", cs, ncl, ncl * cs, cudaGetErrorString(e));
    if (e != cudaSuccess || ncl < 1) { cudaGetLastError(); continue; }
    // DSMEM round trip on one cluster, then on all co-resident clusters at once
    for (int nclu : {1, ncl}) {
      CK(cudaMemset(fail, 0, 4));
      cfg.gridDim = dim3(cs * nclu);
      int iters = 208;
      e = cudaLaunchKernelEx(&cfg, cluster_kernel, iters, out, fail);
      if (e != cudaSuccess) { printf("    launch failed: %s\n", cudaGetErrorString(e)); cudaGetLastError(); continue; }
      CK(cudaDeviceSynchronize());
      long long cyc; int f;
      CK(cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&f, fail, 4, cudaMemcpyDeviceToHost));
      printf("    %3d cluster(s): DSMEM round trip (128 words out to every peer, 128 words back) %6.0f cycles  fail=%d\n", nclu, (double)cyc / (iters - 8), f);
    }
  }
}


// =====================================================================================================================
// D: anatomy of one L2 signalling hop.  CTA 0 writes a flag, G-1 CTAs poll it and answer with one word each, CTA 0 polls the
// answers.  cycles / iteration = two hops.  st_kind: 0 st.relaxed.gpu, 1 st.release.gpu, 2 atom.exch, 3 st.volatile,
// 4 st.relaxed.gpu + fence.acq_rel.gpu.  ld_kind: 0 ld.relaxed.gpu, 1 ld.acquire.gpu, 2 ld.volatile, 3 atom.add 0,
// 4 ld.relaxed.gpu with 64 ns back-off.  ack_stride: words between the answers of consecutive CTAs (1 = packed, 16 = own line)
// =====================================================================================================================
__device__ __forceinline__ void st_kind_u64(unsigned long long* p, unsigned long long v, int k) {
  if (k == 0) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
  else if (k == 1) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
  else if (k == 2) atomicExch(p, v);
  else if (k == 3) *reinterpret_cast<volatile unsigned long long*>(p) = v;
  else { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
}
__device__ __forceinline__ unsigned long long ld_kind_u64(unsigned long long* p, int k) {
  unsigned long long v;
  if (k == 0 || k == 4) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  else if (k == 1) asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  else if (k == 2) v = *reinterpret_cast<volatile unsigned long long*>(p);
  else v = atomicAdd(p, 0ull);
  if (k == 4) __nanosleep(64);
  return v;
}
__global__ void __launch_bounds__(128, 1) hop_kernel(int st_kind, int ld_kind, int ack_stride, int iters, unsigned long long* flag, unsigned long long* ack,
                                                      long long* out, int* fail) {
  extern __shared__ unsigned char smem_raw[];
  (void)smem_raw;
  const int G = gridDim.x, lane = threadIdx.x & 31;
  if (threadIdx.x >= 32) return;
  if (blockIdx.x == 0) {
    long long t0 = 0;
    for (int it = 1; it <= iters; it++) {
      if (it == 9) t0 = clock64();
      if (lane == 0) st_kind_u64(flag, (unsigned long long)it, st_kind);
      uint32_t spin = 0;
      bool dead = false;
      while (!dead) {
        bool ok = true;
        for (int w = 1 + lane; w < G; w += 32) ok = ok && (ld_kind_u64(ack + (size_t)w * ack_stride, ld_kind) == (unsigned long long)it);
        if (__all_sync(0xffffffffu, ok)) break;
        if (++spin > kSpin) { dead = true; *fail = 1; }
      }
      if (dead) break;
    }
    if (lane == 0) out[0] = clock64() - t0;
  } else {
    for (int it = 1; it <= iters; it++) {
      if (lane == 0) {
        uint32_t spin = 0;
        while (ld_kind_u64(flag, ld_kind) != (unsigned long long)it) if (++spin > kSpin) { *fail = 2; break; }
        st_kind_u64(ack + (size_t)blockIdx.x * ack_stride, (unsigned long long)it, st_kind);
      }
      __syncwarp();
      if (*reinterpret_cast<volatile int*>(fail)) break;
    }
  }
}
static void run_D(int sms) {
  unsigned long long *flag, *ack; long long* out; int* fail;
  CK(cudaMalloc(&flag, 4096)); CK(cudaMalloc(&ack, 8 * 16 * 160)); CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&fail, 4));
  const size_t smem = 120 * 1024;
  CK(cudaFuncSetAttribute(hop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  printf("D: flag -> G-1 pollers -> one answer word each -> CTA 0; cycles per iteration (= two L2 hops)\n");
  const char* stn[] = {"st.relaxed", "st.release", "atom.exch", "st.volatile", "st.relaxed+fence"};
  const char* ldn[] = {"ld.relaxed", "ld.acquire", "ld.volatile", "atom.add0", "ld.relaxed+sleep"};
  for (int G : {2, 4, 8, 16, 32, 74, 148}) {
    if (G > sms) continue;
    for (int stride : {1, 16})
      for (int sk = 0; sk < 5; sk++)
        for (int lk = 0; lk < 5; lk++) {
          if (stride == 16 && !(sk == 0 && lk == 0) && !(sk == 2 && lk == 3)) continue;
          if (G != 2 && G != 16 && G != 148 && !((sk == 0 && lk == 0) || (sk == 2 && lk == 3))) continue;
          CK(cudaMemset(flag, 0, 4096)); CK(cudaMemset(ack, 0, 8 * 16 * 160)); CK(cudaMemset(fail, 0, 4));
          int iters = 508, a_sk = sk, a_lk = lk, a_stride = stride;
          void* args[] = {&a_sk, &a_lk, &a_stride, &iters, &flag, &ack, &out, &fail};
          CK(cudaLaunchCooperativeKernel((void*)hop_kernel, dim3(G), dim3(128), args, smem, 0));
          CK(cudaDeviceSynchronize());
          long long cyc; int f;
          CK(cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&f, fail, 4, cudaMemcpyDeviceToHost));
          printf("  G=%3d ack_stride=%2d %-17s %-17s %7.0f cycles/iter fail=%d\n", G, stride, stn[sk], ldn[lk], (double)cyc / (iters - 8), f);
        }
  }
}

int main(int argc, char** argv) {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
  const char* which = argc > 1 ? argv[1] : "ABC";
  if (strchr(which, 'A')) run_A();
  if (strchr(which, 'B')) run_B(prop.multiProcessorCount);
  if (strchr(which, 'C')) run_C();
  if (strchr(which, 'D')) run_D(prop.multiProcessorCount);
  return 0;
}
