// Host-side zeroing for the sparse device -> host route (eri_api.cu): the zeros of a run of unflagged chunks are
// written with non-temporal stores, so that a cache line is written to DRAM once instead of being read for ownership
// first (the plain memset of these runs, which are mostly below glibc's non-temporal threshold, moved twice the bytes
// and was the slowest part of the host-buffer call: profiles/r2_notes.md section 7).
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <immintrin.h>

namespace myqc {

__attribute__((target("avx2"))) static void zero_lines_avx2(double* p, size_t nlines) {
    const __m256i z = _mm256_setzero_si256();
    for (size_t i = 0; i < nlines; ++i, p += 8) {
        _mm256_stream_si256(reinterpret_cast<__m256i*>(p), z);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(p + 4), z);
    }
}

static void zero_lines_sse2(double* p, size_t nlines) {
    const __m128i z = _mm_setzero_si128();
    for (size_t i = 0; i < nlines; ++i, p += 8) {
        _mm_stream_si128(reinterpret_cast<__m128i*>(p), z);
        _mm_stream_si128(reinterpret_cast<__m128i*>(p + 2), z);
        _mm_stream_si128(reinterpret_cast<__m128i*>(p + 4), z);
        _mm_stream_si128(reinterpret_cast<__m128i*>(p + 6), z);
    }
}

// n doubles at p (8-byte aligned) become +0.0; call host_zero_fence() once per thread after the last run
void host_zero_stream(double* p, size_t n) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (n < 32) { std::memset(p, 0, n * sizeof(double)); return; }
    while (n && (reinterpret_cast<uintptr_t>(p) & 63)) { *p++ = 0.0; --n; }
    const size_t nlines = n / 8;
    if (avx2) zero_lines_avx2(p, nlines); else zero_lines_sse2(p, nlines);
    p += nlines * 8;
    n -= nlines * 8;
    while (n) { *p++ = 0.0; --n; }
}

void host_zero_fence() { _mm_sfence(); }

}  // namespace myqc

// C-ABI entry (include/myqc_eri.h): n doubles at p become +0.0 with the streaming-store routine of the sparse
// device -> host route, fenced before it returns.  Exposed so that the host-side tests can check heads, tails and
// alignments without a GPU.
extern "C" void myqc_host_zero(double* p, int64_t n) {
    if (!p || n <= 0) return;
    myqc::host_zero_stream(p, (size_t)n);
    myqc::host_zero_fence();
}
