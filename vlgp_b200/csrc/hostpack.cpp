// Host-side packing loops of the upload path (plain C++, no CUDA): spike counts arrive as float64 arrays (the
// reference promotes y to float64, vlgp/core.py:60) and are stored in HBM as uint8 when every entry is an integer in
// [0, 255].  The conversion + exactness check reads 8 bytes and writes 1 per entry and runs in the
// host threads of the pinned upload pipeline (capi.cu: vlgp_trials_set_y_parts); it is the largest host cost of a
// vem() call with host buffers (204 MB of float64 per call at 256 trials x 1000 bins x 100 neurons), so it gets an AVX2
// body chosen at run time: 2.8 -> 5.8 GB/s of float64 per thread.
#include <stdint.h>

#include "../../include/vlgp_b200.h"

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define VLGP_HAVE_AVX2_DISPATCH 1
#endif

namespace {

bool f64_to_u8_scalar(const double *src, unsigned char *dst, int64_t cnt) {
    bool ok = true;
    for (int64_t k = 0; k < cnt; ++k) {
        const double v = src[k];
        const bool in = (v >= 0.0) & (v <= 255.0);                     // also false for NaN
        const unsigned char c = in ? (unsigned char)(int)v : (unsigned char)0;
        dst[k] = c;
        ok &= in & ((double)c == v);                                   // exact only for integer counts
    }
    return ok;
}

#ifdef VLGP_HAVE_AVX2_DISPATCH
__attribute__((target("avx2"))) bool f64_to_u8_avx2(const double *src, unsigned char *dst, int64_t cnt) {
    int64_t k = 0;
    __m256d bad = _mm256_setzero_pd();
    const __m256d lo = _mm256_set1_pd(0.0), hi = _mm256_set1_pd(255.0);
    const __m256d ones = _mm256_castsi256_pd(_mm256_set1_epi64x(-1));
    for (; k + 16 <= cnt; k += 16) {
        __m128i q[4];
        for (int j = 0; j < 4; ++j) {
            const __m256d v = _mm256_loadu_pd(src + k + 4 * j);
            const __m128i i32 = _mm256_cvttpd_epi32(v);               // truncates; out of range / NaN -> INT_MIN
            const __m256d back = _mm256_cvtepi32_pd(i32);
            // an exact count: v equals its truncation and lies in [0, 255] (ordered compares: NaN fails all three)
            const __m256d good = _mm256_and_pd(
                _mm256_cmp_pd(v, back, _CMP_EQ_OQ),
                _mm256_and_pd(_mm256_cmp_pd(v, lo, _CMP_GE_OQ), _mm256_cmp_pd(v, hi, _CMP_LE_OQ)));
            bad = _mm256_or_pd(bad, _mm256_andnot_pd(good, ones));
            q[j] = i32;
        }
        // saturating packs: what they store for an inexact entry is irrelevant, the caller discards the whole buffer
        const __m128i w0 = _mm_packus_epi32(q[0], q[1]), w1 = _mm_packus_epi32(q[2], q[3]);
        _mm_storeu_si128((__m128i *)(dst + k), _mm_packus_epi16(w0, w1));
    }
    bool ok = _mm256_movemask_pd(bad) == 0;
    if (k < cnt) ok &= f64_to_u8_scalar(src + k, dst + k, cnt - k);
    return ok;
}
#endif

}   // namespace

// dst[k] = (uint8) src[k] for k < cnt; returns 1 when every src[k] is an integer in [0, 255] (then dst is exact), else 0
// (dst then holds unspecified bytes).  Thread-safe; called concurrently on disjoint ranges.
int vlgp_host_f64_to_u8(const double *src, unsigned char *dst, int64_t cnt) {
#ifdef VLGP_HAVE_AVX2_DISPATCH
    static const int have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) return f64_to_u8_avx2(src, dst, cnt) ? 1 : 0;
#endif
    return f64_to_u8_scalar(src, dst, cnt) ? 1 : 0;
}

// Which body vlgp_host_f64_to_u8 runs on this machine: 1 = AVX2, 0 = scalar (tests / diagnostics).
int vlgp_host_pack_isa(void) {
#ifdef VLGP_HAVE_AVX2_DISPATCH
    return __builtin_cpu_supports("avx2") ? 1 : 0;
#else
    return 0;
#endif
}
