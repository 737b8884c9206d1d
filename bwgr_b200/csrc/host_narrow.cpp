// host_narrow.cpp -- exact narrowing of R's double genotype matrix (what emRR(y, gen) receives, R/RcppExports.R) to int8 codes
// on the host cores, one column at a time.  Pure host code: the vector paths are selected at run time (the build box and the GPU
// box need not have the same CPU).  A value that is not an integer code in [lo, hi] sets the returned flag; the caller rejects
// the matrix -- nothing is rounded silently.
#include <immintrin.h>

#include <cmath>
#include <climits>
#include <cstdint>

namespace bwgr {

namespace {

int narrow_plain(const double* src, int64_t n, int8_t* out, int lo, int hi) {
  int flag = 0;
  for (int64_t i = 0; i < n; i++) {
    const double v = src[i];
    const int iv = (v >= -2147483648.0 && v < 2147483648.0) ? (int)v : INT32_MIN;  // NaN -> INT32_MIN -> out of range
    flag |= !((double)iv == v) | (iv < lo) | (iv > hi);
    out[i] = (int8_t)iv;
  }
  return flag;
}

int shift_plain(const double* src, int64_t n, int8_t* out, double shift, int lo, int hi) {
  int flag = 0;
  for (int64_t i = 0; i < n; i++) {
    const double t = src[i] - shift;
    const double r = std::nearbyint(t);
    flag |= !(std::fabs(t - r) <= 1e-4) | !(r >= lo) | !(r <= hi);
    out[i] = (int8_t)(int)((r >= lo && r <= hi) ? r : 0.0);
  }
  return flag;
}

double min_plain(const double* src, int64_t n) {
  double mn = 1e300;
  for (int64_t i = 0; i < n; i++) mn = src[i] < mn ? src[i] : mn;
  return mn;
}

__attribute__((target("avx512f,avx512bw,avx512vl,avx512dq"))) int narrow_512(const double* src, int64_t n, int8_t* out, int lo, int hi) {
  const __m512i vlo = _mm512_set1_epi32(lo), vhi = _mm512_set1_epi32(hi);
  unsigned bad = 0;
  int64_t i = 0;
  for (; i + 16 <= n; i += 16) {
    const __m512d a = _mm512_loadu_pd(src + i), b = _mm512_loadu_pd(src + i + 8);
    const __m256i ia = _mm512_cvttpd_epi32(a), ib = _mm512_cvttpd_epi32(b);  // out of int32 range / NaN -> INT32_MIN
    bad |= _mm512_cmp_pd_mask(_mm512_cvtepi32_pd(ia), a, _CMP_NEQ_UQ) | _mm512_cmp_pd_mask(_mm512_cvtepi32_pd(ib), b, _CMP_NEQ_UQ);
    const __m512i v = _mm512_inserti64x4(_mm512_castsi256_si512(ia), ib, 1);
    bad |= _mm512_cmp_epi32_mask(v, vlo, _MM_CMPINT_LT) | _mm512_cmp_epi32_mask(v, vhi, _MM_CMPINT_NLE);
    _mm_storeu_si128(reinterpret_cast<__m128i*>(out + i), _mm512_cvtepi32_epi8(v));
  }
  return (bad != 0) | narrow_plain(src + i, n - i, out + i, lo, hi);
}

__attribute__((target("avx512f,avx512bw,avx512vl,avx512dq"))) int shift_512(const double* src, int64_t n, int8_t* out, double shift, int lo, int hi) {
  const __m512i vlo = _mm512_set1_epi32(lo), vhi = _mm512_set1_epi32(hi);
  const __m512d vs = _mm512_set1_pd(shift), tol = _mm512_set1_pd(1e-4);
  unsigned bad = 0;
  int64_t i = 0;
  for (; i + 16 <= n; i += 16) {
    const __m512d a = _mm512_sub_pd(_mm512_loadu_pd(src + i), vs), b = _mm512_sub_pd(_mm512_loadu_pd(src + i + 8), vs);
    const __m512d ra = _mm512_roundscale_pd(a, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC), rb = _mm512_roundscale_pd(b, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC);
    bad |= _mm512_cmp_pd_mask(_mm512_abs_pd(_mm512_sub_pd(a, ra)), tol, _CMP_NLE_UQ) | _mm512_cmp_pd_mask(_mm512_abs_pd(_mm512_sub_pd(b, rb)), tol, _CMP_NLE_UQ);
    const __m512i v = _mm512_inserti64x4(_mm512_castsi256_si512(_mm512_cvtpd_epi32(ra)), _mm512_cvtpd_epi32(rb), 1);
    bad |= _mm512_cmp_epi32_mask(v, vlo, _MM_CMPINT_LT) | _mm512_cmp_epi32_mask(v, vhi, _MM_CMPINT_NLE);
    _mm_storeu_si128(reinterpret_cast<__m128i*>(out + i), _mm512_cvtepi32_epi8(v));
  }
  return (bad != 0) | shift_plain(src + i, n - i, out + i, shift, lo, hi);
}

__attribute__((target("avx512f,avx512bw,avx512vl,avx512dq"))) double min_512(const double* src, int64_t n) {
  __m512d m = _mm512_set1_pd(1e300);
  int64_t i = 0;
  for (; i + 8 <= n; i += 8) m = _mm512_min_pd(_mm512_loadu_pd(src + i), m);  // min_pd returns the second operand when one is NaN
  const double t = min_plain(src + i, n - i);
  const double v = _mm512_reduce_min_pd(m);
  return t < v ? t : v;
}

__attribute__((target("avx2"))) int narrow_256(const double* src, int64_t n, int8_t* out, int lo, int hi) {
  const __m128i vlo = _mm_set1_epi32(lo), vhi = _mm_set1_epi32(hi);
  int bad = 0;
  int64_t i = 0;
  for (; i + 4 <= n; i += 4) {
    const __m256d a = _mm256_loadu_pd(src + i);
    const __m128i ia = _mm256_cvttpd_epi32(a);
    bad |= _mm256_movemask_pd(_mm256_cmp_pd(_mm256_cvtepi32_pd(ia), a, _CMP_NEQ_UQ));
    bad |= _mm_movemask_epi8(_mm_or_si128(_mm_cmplt_epi32(ia, vlo), _mm_cmpgt_epi32(ia, vhi)));
    const __m128i w = _mm_packs_epi32(ia, ia);  // in range (checked above) -> no saturation
    const int32_t four = _mm_cvtsi128_si32(_mm_packs_epi16(w, w));
    __builtin_memcpy(out + i, &four, 4);
  }
  return (bad != 0) | narrow_plain(src + i, n - i, out + i, lo, hi);
}

int cpu_level() {
  static const int level = [] {
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx512dq")) return 2;
    if (__builtin_cpu_supports("avx2")) return 1;
    return 0;
  }();
  return level;
}

}  // namespace

// int8 code of every value of a column, exactly; nonzero = some value is not an integer code in [lo, hi]
int narrow_column(const double* src, int64_t n, int8_t* out, int lo, int hi) {
  const int l = cpu_level();
  return l == 2 ? narrow_512(src, n, out, lo, hi) : l == 1 ? narrow_256(src, n, out, lo, hi) : narrow_plain(src, n, out, lo, hi);
}

// the same for a column stored as (integer code + one constant): codes of src - shift, each within 1e-4 of an integer
int narrow_column_shifted(const double* src, int64_t n, int8_t* out, double shift, int lo, int hi) {
  return cpu_level() == 2 ? shift_512(src, n, out, shift, lo, hi) : shift_plain(src, n, out, shift, lo, hi);
}

double column_min(const double* src, int64_t n) { return cpu_level() == 2 ? min_512(src, n) : min_plain(src, n); }

}  // namespace bwgr

// Test hook (tests/test_host_logic.py, no GPU needed): one column through a chosen code path.  level: -1 = the one the loader uses on
// this CPU, 0 plain, 1 AVX2, 2 AVX-512 (refused with -1 if this CPU lacks it).  Returns the "not an integer code in range" flag.
extern "C" __attribute__((visibility("default"))) int bwgr_debug_narrow(const double* src, int64_t n, int8_t* out, int use_shift, double shift, int lo, int hi, int level,
                                                                         double* min_out) {
  using namespace bwgr;
  if (level > cpu_level()) return -1;
  if (level < 0) level = cpu_level();
  if (min_out) *min_out = level == 2 ? min_512(src, n) : min_plain(src, n);
  if (use_shift) return level == 2 ? shift_512(src, n, out, shift, lo, hi) : shift_plain(src, n, out, shift, lo, hi);
  return level == 2 ? narrow_512(src, n, out, lo, hi) : level == 1 ? narrow_256(src, n, out, lo, hi) : narrow_plain(src, n, out, lo, hi);
}
