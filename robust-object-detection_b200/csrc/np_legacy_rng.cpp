// np_legacy_rng.cpp -- the noise field of compat mode, drawn on the host exactly as NumPy's legacy generator does.
//
// Reference: scripts/augmentations.py:31  `np.random.normal(0, sigma, img.shape).astype(np.float32)`.  That call is
// the reference's own bottleneck (104 of the 137 ms of apply_noise on a 1360x765 frame): NumPy's global RandomState
// (MT19937) feeding the legacy polar Gaussian (numpy/random/src/legacy/legacy-distributions.c legacy_gauss,
// numpy/random/src/mt19937/mt19937.h mt19937_next_double), one scalar at a time.  In compat mode the GPU kernel
// consumes that field bit for bit, so the field has to be this exact stream.  This file regenerates it faster without
// changing a bit: the MT19937 word stream is produced sequentially (it is a linear recurrence), the polar method's
// rejection test, log, sqrt and division -- independent per candidate pair -- run on all host threads, and the
// generator state handed back (key, pos, has_gauss, cached gaussian) is exactly what NumPy's would be after the call,
// so any later np.random use continues on the same stream.
//
// Bit-exactness rests on: identical integer stream; (a * 2^26 + b) / 2^53 and 2x - 1 exact in double; x1*x1 + x2*x2,
// the division and sqrt are IEEE operations (this file is built with -ffp-contract=off: no FMA); log() is the same libm
// function NumPy calls.  tests/test_host_logic.py checks fields, odd counts, cached-gaussian hand-over and the
// continued stream against np.random itself.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "../../include/rod_b200.h"

namespace {

constexpr int kN = 624, kM = 397;
constexpr uint32_t kMatrixA = 0x9908b0dfu, kUpper = 0x80000000u, kLower = 0x7fffffffu;

inline void mt_regenerate(uint32_t* mt) {
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
inline uint32_t untemper(uint32_t y) {  // inverse of temper(): recovers the raw state word
    y ^= y >> 18;
    y ^= (y << 15) & 0xefc60000u;
    uint32_t t = y;
    for (int i = 0; i < 4; ++i) t = y ^ ((t << 7) & 0x9d2c5680u);
    y = t;
    for (int i = 0; i < 2; ++i) t = y ^ (t >> 11);
    return t;
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
inline Candidate candidate(const uint32_t* w) {
    Candidate c;
    c.x1 = 2.0 * word_pair_to_double(w[0], w[1]) - 1.0;
    c.x2 = 2.0 * word_pair_to_double(w[2], w[3]) - 1.0;
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
    // tempered stream: leftover of the current block, then whole blocks.  Kept per calling thread between calls: a fresh
    // 32 MB buffer per frame costs more in page faults than generating the words
    static thread_local std::vector<uint32_t> tl_words;
    std::vector<uint32_t>& words = tl_words;  // (a local reference: the worker lambdas must see THIS thread's buffer)
    std::vector<uint64_t> counts;
    while (done < n) {
        const uint64_t need_pairs = (n - done + 1) / 2;
        // candidates to draw this round: expectation 4/pi per accepted pair, plus slack; bounded per round
        uint64_t cand = (uint64_t)((double)need_pairs * 1.2740) + 64;
        cand = std::min<uint64_t>(cand, 1ull << 24);
        const uint64_t want_words = 4 * cand;
        // the stream from the current position: rest of the present block, then as many fresh blocks as needed
        words.clear();
        words.reserve(want_words + kN);
        const int start_pos = *pos;
        for (int i = start_pos; i < kN; ++i) words.push_back(temper(key[i]));
        std::vector<uint32_t> block(key, key + kN);
        while (words.size() < want_words) {
            mt_regenerate(block.data());
            const size_t o = words.size();
            words.resize(o + kN);
            for (int i = 0; i < kN; ++i) words[o + i] = temper(block[i]);
        }
        cand = words.size() / 4;
        // pass 1: acceptance counts per range
        int used_threads = 1;
        counts.assign((size_t)threads + 1, 0);
        parallel_ranges(cand, threads, [&](int t, uint64_t lo, uint64_t hi) {
            uint64_t c = 0;
            for (uint64_t i = lo; i < hi; ++i) c += candidate(&words[4 * i]).ok ? 1 : 0;
            counts[(size_t)t + 1] = c;
        });
        used_threads = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)threads, (cand + 4095) / 4096));
        for (int t = 0; t < used_threads; ++t) counts[(size_t)t + 1] += counts[(size_t)t];
        const uint64_t accepted = counts[(size_t)used_threads];
        const uint64_t take_pairs = std::min(accepted, need_pairs);
        // pass 2: outputs of the first take_pairs accepted candidates; remember the index of the last one
        std::vector<uint64_t> last_idx((size_t)used_threads, 0);
        double tail_cached = 0.0;
        bool tail_has = false;
        parallel_ranges(cand, threads, [&](int t, uint64_t lo, uint64_t hi) {
            uint64_t rank = counts[(size_t)t];
            for (uint64_t i = lo; i < hi && rank < take_pairs; ++i) {
                const Candidate c = candidate(&words[4 * i]);
                if (!c.ok) continue;
                const double f = sqrt(-2.0 * log(c.r2) / c.r2);
                const double g_first = f * c.x2, g_second = f * c.x1;  // returned now / kept for the next call
                const uint64_t o = done + 2 * rank;
                out[o] = (float)(0.0 + sigma * g_first);
                if (o + 1 < n) out[o + 1] = (float)(0.0 + sigma * g_second);
                else { tail_cached = g_second; tail_has = true; }  // only the very last pair of an odd count
                ++rank;
                last_idx[(size_t)t] = i + 1;  // candidates consumed up to here (exclusive)
            }
        });
        uint64_t consumed_cand = cand;  // all of them when this round did not reach the target
        if (take_pairs == need_pairs) {
            consumed_cand = 0;
            for (int t = 0; t < used_threads; ++t) consumed_cand = std::max(consumed_cand, last_idx[(size_t)t]);
        }
        done = std::min<uint64_t>(n, done + 2 * take_pairs);
        if (tail_has) { *has_gauss = 1; *cached = tail_cached; }
        // generator state after consuming 4 * consumed_cand words from start_pos
        const uint64_t consumed_words = 4 * consumed_cand;
        const uint64_t in_first = (uint64_t)(kN - start_pos);
        if (consumed_words <= in_first) {
            *pos = start_pos + (int)consumed_words;  // still inside the block the call started in: key unchanged
        } else {
            const uint64_t beyond = consumed_words - in_first;         // words taken from fresh blocks
            const uint64_t blk = (beyond - 1) / kN;                    // 0-based fresh block holding the last word
            const uint32_t* bw = &words[in_first + blk * kN];          // its tempered words -> raw state
            for (int i = 0; i < kN; ++i) key[i] = untemper(bw[i]);
            *pos = (int)(beyond - blk * kN);
        }
    }
    return ROD_OK;
}
