// rgb_expand.cpp -- see rgb_expand.h
#include "rgb_expand.h"

#include <string.h>
#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define J2K_X86 1
#endif

namespace {

void expand_row_scalar(const uint8_t *s, uint8_t *d, uint32_t n)
{
    for (uint32_t i = 0; i < n; i++) { d[4 * i] = s[3 * i]; d[4 * i + 1] = s[3 * i + 1]; d[4 * i + 2] = s[3 * i + 2]; d[4 * i + 3] = 255; }
}

#ifdef J2K_X86
// 4 pixels per step: 12 source bytes (read as 16, so the last steps of a row go through the scalar loop) -> 16 bytes
__attribute__((target("ssse3"))) void expand_row_ssse3(const uint8_t *s, uint8_t *d, uint32_t n)
{
    const __m128i shuf = _mm_setr_epi8(0, 1, 2, -1, 3, 4, 5, -1, 6, 7, 8, -1, 9, 10, 11, -1);
    const __m128i alpha = _mm_set1_epi32((int)0xFF000000u);
    uint32_t i = 0;
    for (; i + 8 <= n; i += 4) {
        const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + 3 * i));
        _mm_storeu_si128(reinterpret_cast<__m128i *>(d + 4 * i), _mm_or_si128(_mm_shuffle_epi8(v, shuf), alpha));
    }
    expand_row_scalar(s + 3 * i, d + 4 * i, n - i);
}

// 8 pixels per step: two 12-byte groups, one per 128-bit lane
__attribute__((target("avx2"))) void expand_row_avx2(const uint8_t *s, uint8_t *d, uint32_t n)
{
    const __m256i shuf = _mm256_setr_epi8(0, 1, 2, -1, 3, 4, 5, -1, 6, 7, 8, -1, 9, 10, 11, -1,
                                          0, 1, 2, -1, 3, 4, 5, -1, 6, 7, 8, -1, 9, 10, 11, -1);
    const __m256i alpha = _mm256_set1_epi32((int)0xFF000000u);
    uint32_t i = 0;
    for (; i + 12 <= n; i += 8) {
        const __m128i lo = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + 3 * i));
        const __m128i hi = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + 3 * i + 12));
        const __m256i v = _mm256_inserti128_si256(_mm256_castsi128_si256(lo), hi, 1);
        _mm256_storeu_si256(reinterpret_cast<__m256i *>(d + 4 * i), _mm256_or_si256(_mm256_shuffle_epi8(v, shuf), alpha));
    }
    expand_row_scalar(s + 3 * i, d + 4 * i, n - i);
}
#endif

typedef void (*row_fn)(const uint8_t *, uint8_t *, uint32_t);
row_fn pick_row_fn()
{
#ifdef J2K_X86
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx2")) return expand_row_avx2;
    if (__builtin_cpu_supports("ssse3")) return expand_row_ssse3;
#endif
    return expand_row_scalar;
}

}  // namespace

void j2k_expand_rgb24(const J2kExpandTask &t)
{
    static const row_fn fn = pick_row_fn();
    for (uint32_t y = 0; y < t.rows; y++) fn(t.src + (uint64_t)y * t.sstride, t.dst + (uint64_t)y * t.dstride, t.width);
}

void J2kExpandPool::start(unsigned threads)
{
    std::lock_guard<std::mutex> g(mu_);
    if (!th_.empty() || threads == 0) return;
    quit_ = false;
    for (unsigned i = 0; i < threads; i++) th_.emplace_back([this] { run(); });
}

void J2kExpandPool::submit(const J2kExpandTask &t)
{
    const uint32_t band = 64;
    {
        std::lock_guard<std::mutex> g(mu_);
        for (uint32_t y = 0; y < t.rows; y += band) {
            J2kExpandTask b = t;
            b.src = t.src + (uint64_t)y * t.sstride; b.dst = t.dst + (uint64_t)y * t.dstride;
            b.rows = t.rows - y < band ? t.rows - y : band;
            q_.push_back(b);
            pending_++;
        }
    }
    cv_.notify_all();
}

void J2kExpandPool::run()
{
    for (;;) {
        J2kExpandTask t;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [this] { return quit_ || !q_.empty(); });
            if (q_.empty()) return;                      // quit_
            t = q_.front();
            q_.pop_front();
        }
        j2k_expand_rgb24(t);
        {
            std::lock_guard<std::mutex> g(mu_);
            if (--pending_ == 0) idle_.notify_all();
        }
    }
}

void J2kExpandPool::wait_idle()
{
    std::unique_lock<std::mutex> lk(mu_);
    if (th_.empty()) {                                   // no workers: the caller's thread does the work
        while (!q_.empty()) {
            const J2kExpandTask t = q_.front();
            q_.pop_front();
            lk.unlock();
            j2k_expand_rgb24(t);
            lk.lock();
            pending_--;
        }
        return;
    }
    idle_.wait(lk, [this] { return pending_ == 0; });
}

void J2kExpandPool::stop()
{
    {
        std::lock_guard<std::mutex> g(mu_);
        quit_ = true;
    }
    cv_.notify_all();
    for (auto &t : th_) t.join();
    th_.clear();
}
