// np_legacy_rng.cpp -- the noise field of compat mode, drawn on the host exactly as NumPy's legacy generator does.
//
// Reference: scripts/augmentations.py:31  `np.random.normal(0, sigma, img.shape).astype(np.float32)`.  That call is
// the reference's own bottleneck (104 of the 137 ms of apply_noise on a 1360x765 frame): NumPy's global RandomState
// (MT19937) feeding the legacy polar Gaussian (numpy/random/src/legacy/legacy-distributions.c legacy_gauss,
// numpy/random/src/mt19937/mt19937.h mt19937_next_double), one scalar at a time.  In compat mode the GPU kernel
// consumes that field bit for bit, so the field has to be this exact stream.  This file regenerates it faster without
// changing a bit: the calling thread produces the MT19937 word stream (a linear recurrence: sequential) in 0.5 MB chunks
// while the other host threads already run the polar method on the chunks that are ready -- rejection test, log, sqrt
// and division are independent per candidate pair -- each chunk's accepted pairs are then compacted into the output
// in order, and the generator state handed back (key, pos, has_gauss, cached gaussian) is exactly what NumPy's would
// be after the call, so any later np.random use continues on the same stream.
//
// Bit-exactness rests on: identical integer stream; (a * 2^26 + b) / 2^53 and 2x - 1 exact in double; x1*x1 + x2*x2,
// the division and sqrt are IEEE operations (this file is built with -ffp-contract=off: no FMA); log() is the same libm
// function NumPy calls.  tests/test_host_logic.py checks fields, odd counts, cached-gaussian hand-over and the
// continued stream against np.random itself.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <new>
#include <system_error>
#include <thread>
#include <vector>

#include "../../include/rod_b200.h"

namespace {

constexpr int kN = 624, kM = 397;
constexpr uint32_t kMatrixA = 0x9908b0dfu, kUpper = 0x80000000u, kLower = 0x7fffffffu;

// (cloned for AVX2 with run-time dispatch: both loops vectorise, dependence distances are 227 and 397 words)
__attribute__((target_clones("avx2", "default"))) void mt_regenerate(uint32_t* mt) {
    int i = 0;
    uint32_t y;
    for (; i < kN - kM; ++i) {
        y = (mt[i] & kUpper) | (mt[i + 1] & kLower);
        mt[i] = mt[i + kM] ^ (y >> 1) ^ ((0u - (y & 1u)) & kMatrixA);
    }
    for (; i < kN - 1; ++i) {
        y = (mt[i] & kUpper) | (mt[i + 1] & kLower);
        mt[i] = mt[i + (kM - kN)] ^ (y >> 1) ^ ((0u - (y & 1u)) & kMatrixA);
    }
    y = (mt[kN - 1] & kUpper) | (mt[0] & kLower);
    mt[kN - 1] = mt[kM - 1] ^ (y >> 1) ^ ((0u - (y & 1u)) & kMatrixA);
}
inline uint32_t temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}
// mt19937_next_double: 53-bit double in [0, 1) from two words
inline double word_pair_to_double(uint32_t w0, uint32_t w1) {
    const int32_t a = (int32_t)(w0 >> 5), b = (int32_t)(w1 >> 6);
    return (a * 67108864.0 + b) / 9007199254740992.0;
}
struct Candidate {
    double x1, x2, r2;
    bool ok;
};
inline Candidate candidate(const uint32_t* w) {  // w: four RAW state words (tempered here, on the worker threads)
    Candidate c;
    c.x1 = 2.0 * word_pair_to_double(temper(w[0]), temper(w[1])) - 1.0;
    c.x2 = 2.0 * word_pair_to_double(temper(w[2]), temper(w[3])) - 1.0;
    c.r2 = c.x1 * c.x1 + c.x2 * c.x2;
    c.ok = !(c.r2 >= 1.0 || c.r2 == 0.0);
    return c;
}

template <typename F>
void parallel_ranges(uint64_t n, int threads, F fn) {
    threads = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)threads, (n + 4095) / 4096));
    if (threads == 1) { fn(0, 0, n); return; }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(fn, t, n * t / threads, n * (t + 1) / threads);
    for (auto& th : pool) th.join();
}

}  // namespace

// RandomState state in, field out, state after the call out.  key: 624 words (np.random.get_state()[1]), *pos the
// index of the next word (624 = regenerate first), *has_gauss / *cached the legacy cached Gaussian.
extern "C" ROD_API int rod_numpy_legacy_normal_f32(uint32_t* key, int32_t* pos, int32_t* has_gauss, double* cached,
                                                   double sigma, uint64_t n, float* out, int threads) {
    if (key == nullptr || pos == nullptr || has_gauss == nullptr || cached == nullptr || (n > 0 && out == nullptr))
        return ROD_ERR_INVALID_ARG;
    if (*pos < 0 || *pos > kN) return ROD_ERR_INVALID_ARG;
    if (threads < 1) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    uint64_t done = 0;
    if (n > 0 && *has_gauss) {  // legacy_gauss hands out the cached value first
        out[done++] = (float)(0.0 + sigma * *cached);
        *has_gauss = 0;
        *cached = 0.0;
    }
    // raw MT19937 word stream: leftover of the current block, then whole blocks; and the accepted pairs of every chunk before
    // they are compacted into `out`.  Kept per calling thread between calls: fresh 32 MB buffers per frame cost more in
    // page faults than generating the words
    static thread_local std::vector<uint32_t> tl_words;
    static thread_local std::vector<float> tl_pairs;
    std::vector<uint32_t>& words = tl_words;  // (local references: the worker lambdas must see THIS thread's buffers)
    std::vector<float>& pairs = tl_pairs;
    constexpr uint64_t kChunk = 1u << 15;     // candidates per work item (128 K words, 0.5 MB)
    while (done < n) {
        const uint64_t need_pairs = (n - done + 1) / 2;
        // candidates to draw this round: expectation 4/pi per accepted pair, plus slack; bounded per round
        uint64_t cand = (uint64_t)((double)need_pairs * 1.2740) + 64;
        cand = std::min<uint64_t>(cand, 1ull << 24);
        const int start_pos = *pos;
        const uint64_t in_first = (uint64_t)(kN - start_pos);
        const uint64_t fresh_blocks = (4 * cand > in_first) ? (4 * cand - in_first + kN - 1) / kN : 0;
        const uint64_t total_words = in_first + fresh_blocks * kN;
        cand = total_words / 4;
        const uint64_t n_chunks = (cand + kChunk - 1) / kChunk;
        try {
            if (words.size() < total_words) words.resize(total_words);
            if (pairs.size() < 2 * cand) pairs.resize(2 * cand);
        } catch (const std::bad_alloc&) {
            return ROD_ERR_OOM;  // nothing consumed yet in this round: the generator state is still the caller's
        }
        std::vector<uint32_t> chunk_count((size_t)n_chunks, 0);
        std::atomic<uint64_t> words_ready{0}, next_chunk{0};

        // one work item: the accepted pairs of chunk c, in order, as float32 outputs at pairs[2 * c * kChunk ...]
        auto process_chunk = [&](uint64_t c) {
            const uint64_t lo = c * kChunk, hi = std::min(cand, lo + kChunk);
            float* dstp = &pairs[2 * lo];
            uint32_t cnt = 0;
            for (uint64_t i = lo; i < hi; ++i) {
                const Candidate cd = candidate(&words[4 * i]);
                if (!cd.ok) continue;
                const double f = sqrt(-2.0 * log(cd.r2) / cd.r2);
                dstp[2 * cnt] = (float)(0.0 + sigma * (f * cd.x2));      // returned first
                dstp[2 * cnt + 1] = (float)(0.0 + sigma * (f * cd.x1));  // the "cached" second value
                ++cnt;
            }
            chunk_count[(size_t)c] = cnt;
        };
        auto worker = [&]() {
            for (;;) {
                const uint64_t c = next_chunk.fetch_add(1, std::memory_order_relaxed);
                if (c >= n_chunks) return;
                const uint64_t need_words = 4 * std::min(cand, (c + 1) * kChunk);
                while (words_ready.load(std::memory_order_acquire) < need_words) std::this_thread::yield();
                process_chunk(c);
            }
        };
        // the MT19937 word stream is sequential: this thread produces it while the others already consume it
        const int n_workers = cand < 8192 ? 0 : (int)std::min<uint64_t>((uint64_t)std::max(0, threads - 1), n_chunks);
        std::vector<std::thread> pool;
        for (int t = 0; t < n_workers; ++t) {
            try {
                pool.emplace_back(worker);
            } catch (const std::system_error&) {
                break;  // no more threads to be had: the ones running (and this one) do all chunks
            }
        }
        {
            // raw state words: the rest of the present block, then each fresh block regenerated in place from a copy of
            // its predecessor (the tempering is left to the consumers: this thread is the serial part)
            uint64_t w = 0;
            for (int i = start_pos; i < kN; ++i) words[w++] = key[i];
            const uint32_t* prev = key;
            uint64_t published = 0;
            for (uint64_t b = 0; b < fresh_blocks; ++b) {
                memcpy(&words[w], prev, sizeof(uint32_t) * kN);
                mt_regenerate(&words[w]);
                prev = &words[w];
                w += kN;
                if (w - published >= 4 * kChunk / 2) { words_ready.store(w, std::memory_order_release); published = w; }
            }
            words_ready.store(w, std::memory_order_release);
        }
        worker();  // the producer helps with whatever is left (all of it when threads == 1)
        for (auto& th : pool) th.join();

        // compaction: chunk c's pairs go to out[done + 2 * prefix(c) ...]; the need_pairs-th accepted pair ends the round
        std::vector<uint64_t> prefix((size_t)n_chunks + 1, 0);
        for (uint64_t c = 0; c < n_chunks; ++c) prefix[(size_t)c + 1] = prefix[(size_t)c] + chunk_count[(size_t)c];
        const uint64_t accepted = prefix[(size_t)n_chunks];
        const uint64_t take_pairs = std::min(accepted, need_pairs);
        parallel_ranges(n_chunks, threads, [&](int, uint64_t clo, uint64_t chi) {
            for (uint64_t c = clo; c < chi; ++c) {
                const uint64_t p0 = prefix[(size_t)c];
                if (p0 >= take_pairs) break;
                const uint64_t np = std::min<uint64_t>(chunk_count[(size_t)c], take_pairs - p0);
                const uint64_t o = done + 2 * p0;
                const uint64_t nf = std::min<uint64_t>(2 * np, n - o);  // an odd count drops the very last second value
                memcpy(out + o, &pairs[2 * c * kChunk], nf * sizeof(float));
            }
        });
        uint64_t consumed_cand = cand;  // all of them when this round did not reach the target
        if (take_pairs == need_pairs) {
            // the candidate that produced the last pair: rescan its chunk
            uint64_t c = 0;
            while (prefix[(size_t)c + 1] < need_pairs) ++c;
            uint64_t left = need_pairs - prefix[(size_t)c];
            uint64_t i = c * kChunk;
            Candidate last = candidate(&words[4 * i]);
            for (;; ++i) {
                last = candidate(&words[4 * i]);
                if (last.ok && --left == 0) break;
            }
            consumed_cand = i + 1;
            if (done + 2 * need_pairs > n) {  // odd count: the second value of the last pair stays cached, as a double
                const double f = sqrt(-2.0 * log(last.r2) / last.r2);
                *has_gauss = 1;
                *cached = f * last.x1;
            }
        }
        done = std::min<uint64_t>(n, done + 2 * take_pairs);
        const uint64_t consumed_words = 4 * consumed_cand;
        // generator state after consuming those words from start_pos
        if (consumed_words <= in_first) {
            *pos = start_pos + (int)consumed_words;  // still inside the block the call started in: key unchanged
        } else {
            const uint64_t beyond = consumed_words - in_first;         // words taken from fresh blocks
            const uint64_t blk = (beyond - 1) / kN;                    // 0-based fresh block holding the last word
            memcpy(key, &words[in_first + blk * kN], sizeof(uint32_t) * kN);  // its raw words are the state
            *pos = (int)(beyond - blk * kN);
        }
    }
    return ROD_OK;
}
