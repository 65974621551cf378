// Host-side 2-bit packing variants (the gather of gcg_search, superplus_b200/csrc/host_par.cpp) on the
// box's own cores: GB/s of ASCII read per variant and thread count.  Not part of the product; decides
// which variant gcg_pack_stream should use on a given host.   g++ -O3 -std=c++17 -pthread
#include <immintrin.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <thread>
#include <vector>

static inline void nt8 (uint64_t * dst, uint64_t v) { _mm_stream_si64 ((long long *) dst, (long long) v); }

__attribute__ ((target ("avx2"))) static inline uint64_t pack32_avx2 (const char * s)
{
  const __m256i x = _mm256_loadu_si256 ((const __m256i *) s);
  const __m256i c = _mm256_and_si256 (_mm256_srli_epi16 (x, 1), _mm256_set1_epi8 (3));
  const __m256i p = _mm256_maddubs_epi16 (c, _mm256_set1_epi32 (0x01041040));
  const __m256i q = _mm256_madd_epi16 (p, _mm256_set1_epi16 (1));
  const __m256i ctl = _mm256_setr_epi8 (12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                        12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
  const __m256i r = _mm256_shuffle_epi8 (q, ctl);
  return ((uint64_t) (uint32_t) _mm256_extract_epi32 (r, 0) << 32) | (uint32_t) _mm256_extract_epi32 (r, 4);
}
__attribute__ ((target ("avx2"))) void v_avx2_nt8 (uint64_t * dst, const char * src, size_t n)
{ for (size_t w = 0; (w + 1) * 32 <= n; ++w) nt8 (dst + w, pack32_avx2 (src + w * 32)); }
__attribute__ ((target ("avx2"))) void v_avx2_plain (uint64_t * dst, const char * src, size_t n)
{ for (size_t w = 0; (w + 1) * 32 <= n; ++w) dst[w] = pack32_avx2 (src + w * 32); }
__attribute__ ((target ("avx2"))) void v_avx2_nt8_pf (uint64_t * dst, const char * src, size_t n)
{ for (size_t w = 0; (w + 1) * 32 <= n; ++w) { _mm_prefetch (src + w * 32 + 1024, _MM_HINT_NTA); nt8 (dst + w, pack32_avx2 (src + w * 32)); } }

#define T512 __attribute__ ((target ("avx512f,avx512bw,avx512vbmi,avx512vl")))
T512 static inline __m512i fold64 (const char * s)      // 64 bases -> 16 dwords whose low byte holds 4 bases
{
  const __m512i x = _mm512_loadu_si512 ((const void *) s);
  const __m512i c = _mm512_and_si512 (_mm512_srli_epi16 (x, 1), _mm512_set1_epi8 (3));
  return _mm512_madd_epi16 (_mm512_maddubs_epi16 (c, _mm512_set1_epi32 (0x01041040)), _mm512_set1_epi16 (1));
}
T512 static inline __m128i pack64_512 (const char * s)
{
  const __m512i idx = _mm512_castsi128_si512 (_mm_setr_epi8 (28, 24, 20, 16, 12, 8, 4, 0, 60, 56, 52, 48, 44, 40, 36, 32));
  return _mm512_castsi512_si128 (_mm512_permutexvar_epi8 (idx, fold64 (s)));
}
T512 void v_512_nt16 (uint64_t * dst, const char * src, size_t n)
{ for (size_t w = 0; (w + 2) * 32 <= n; w += 2) _mm_stream_si128 ((__m128i *) (dst + w), pack64_512 (src + w * 32)); }
T512 void v_512_plain (uint64_t * dst, const char * src, size_t n)
{ for (size_t w = 0; (w + 2) * 32 <= n; w += 2) _mm_storeu_si128 ((__m128i *) (dst + w), pack64_512 (src + w * 32)); }
// 256 bases -> one 64-byte line, one full-line non-temporal store (dst 64-byte aligned)
T512 void v_512_nt64 (uint64_t * dst, const char * src, size_t n)
{
  const __m512i idx = _mm512_castsi128_si512 (_mm_setr_epi8 (28, 24, 20, 16, 12, 8, 4, 0, 60, 56, 52, 48, 44, 40, 36, 32));
  for (size_t w = 0; (w + 8) * 32 <= n; w += 8) {
    const __m512i a = _mm512_permutexvar_epi8 (idx, fold64 (src + w * 32)), b = _mm512_permutexvar_epi8 (idx, fold64 (src + w * 32 + 64));
    const __m512i c = _mm512_permutexvar_epi8 (idx, fold64 (src + w * 32 + 128)), d = _mm512_permutexvar_epi8 (idx, fold64 (src + w * 32 + 192));
    __m512i r = _mm512_inserti32x4 (a, _mm512_castsi512_si128 (b), 1);
    r = _mm512_inserti32x4 (r, _mm512_castsi512_si128 (c), 2);
    r = _mm512_inserti32x4 (r, _mm512_castsi512_si128 (d), 3);
    _mm512_stream_si512 ((__m512i *) (dst + w), r);
  }
}
T512 void v_512_nt64_pf (uint64_t * dst, const char * src, size_t n)
{
  const __m512i idx = _mm512_castsi128_si512 (_mm_setr_epi8 (28, 24, 20, 16, 12, 8, 4, 0, 60, 56, 52, 48, 44, 40, 36, 32));
  for (size_t w = 0; (w + 8) * 32 <= n; w += 8) {
    _mm_prefetch (src + w * 32 + 2048, _MM_HINT_NTA); _mm_prefetch (src + w * 32 + 2112, _MM_HINT_NTA);
    _mm_prefetch (src + w * 32 + 2176, _MM_HINT_NTA); _mm_prefetch (src + w * 32 + 2240, _MM_HINT_NTA);
    const __m512i a = _mm512_permutexvar_epi8 (idx, fold64 (src + w * 32)), b = _mm512_permutexvar_epi8 (idx, fold64 (src + w * 32 + 64));
    const __m512i c = _mm512_permutexvar_epi8 (idx, fold64 (src + w * 32 + 128)), d = _mm512_permutexvar_epi8 (idx, fold64 (src + w * 32 + 192));
    __m512i r = _mm512_inserti32x4 (a, _mm512_castsi512_si128 (b), 1);
    r = _mm512_inserti32x4 (r, _mm512_castsi512_si128 (c), 2);
    r = _mm512_inserti32x4 (r, _mm512_castsi512_si128 (d), 3);
    _mm512_stream_si512 ((__m512i *) (dst + w), r);
  }
}
void v_memcpy (uint64_t * dst, const char * src, size_t n) { memcpy (dst, src, n / 4); volatile char sink = 0; for (size_t i = 0; i < n; i += 64) sink += src[i]; (void) sink; }

int main (int argc, char ** argv)
{
  const size_t n = (size_t) (argc > 1 ? atoi (argv[1]) : 512) << 20;
  char * src = (char *) aligned_alloc (4096, n);
  uint64_t * dst = (uint64_t *) aligned_alloc (4096, n / 4), * ref = (uint64_t *) aligned_alloc (4096, n / 4);
  for (size_t i = 0; i < n; ++i) src[i] = "ACGTN"[(i * 2654435761u >> 7) % 5];
  memset (dst, 0, n / 4); memset (ref, 0, n / 4);
  const bool has512 = __builtin_cpu_supports ("avx512vbmi");
  printf ("cores online %u, avx512vbmi %d, %zu MB of ASCII\n", std::thread::hardware_concurrency (), (int) has512, n >> 20);
  v_avx2_nt8 (ref, src, n);
  struct V { const char * name; void (*f) (uint64_t *, const char *, size_t); bool need512; bool check; };
  const V vs[] = {{"avx2 nt8 (now)", v_avx2_nt8, false, true}, {"avx2 plain", v_avx2_plain, false, true}, {"avx2 nt8 + prefetch", v_avx2_nt8_pf, false, true},
                  {"avx512 nt16", v_512_nt16, true, true}, {"avx512 plain", v_512_plain, true, true}, {"avx512 nt64", v_512_nt64, true, true},
                  {"avx512 nt64 + prefetch", v_512_nt64_pf, true, true}, {"read + memcpy/4 (floor)", v_memcpy, false, false}};
  const int threads[] = {1, 2, 4, 8, 16};
  for (const V & v : vs) {
    if (v.need512 && !has512) continue;
    printf ("%-26s", v.name);
    for (int nt : threads) {
      if ((unsigned) nt > std::thread::hardware_concurrency ()) break;
      double best = 0;
      const size_t per = n / nt / 4096 * 4096;
      for (int rep = 0; rep < 4; ++rep) {
        auto t0 = std::chrono::steady_clock::now ();
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back ([&, t] () { v.f (dst + t * per / 32, src + t * per, per); });
        for (auto & x : th) x.join ();
        const double dt = std::chrono::duration<double> (std::chrono::steady_clock::now () - t0).count ();
        best = std::max (best, per * nt / dt / 1e9);
      }
      printf ("  %2dT %6.2f%s", nt, best, v.check && memcmp (dst, ref, per * nt / 4) ? "!" : " ");
    }
    printf ("  GB/s\n");
  }
  return 0;
}
