// np_legacy_rng.cpp -- the noise field of compat mode, drawn on the host exactly as NumPy's legacy generator does.
//
// Reference: scripts/augmentations.py:31  `np.random.normal(0, sigma, img.shape).astype(np.float32)`.  That call is
// the reference's own bottleneck (104 of the 137 ms of apply_noise on a 1360x765 frame): NumPy's global RandomState
// (MT19937) feeding the legacy polar Gaussian (numpy/random/src/legacy/legacy-distributions.c legacy_gauss,
// numpy/random/src/mt19937/mt19937.h mt19937_next_double), one scalar at a time.  In compat mode the GPU kernel
// consumes that field bit for bit, so the field has to be this exact stream.  This file regenerates it faster without
// changing a bit: the MT19937 word stream is a linear recurrence (sequential), so the calling thread runs the recurrence
// alone -- in two 2.5 KB buffers that stay in its L1 cache, 0.3 ns per word -- and only publishes the generator STATE at the
// start of every chunk of 52 blocks; the host threads pick chunks up, regenerate the chunk's words from its state into a
// buffer of their own (130 KB: L2) and run the polar method on them -- rejection test, log, sqrt and division are
// independent per candidate pair.  (The first version had the calling thread write the whole 32 MB word stream of a frame
// to memory for the others to read: 7 of the 9-11 ms of a frame were that thread.)  Each chunk's accepted pairs are then
// compacted into the output in order, and the generator state handed back (key, pos, has_gauss, cached gaussian) is exactly
// what NumPy's would be after the call, so any later np.random use continues on the same stream.
//
// Bit-exactness rests on: identical integer stream; (a * 2^26 + b) / 2^53 and 2x - 1 exact in double; x1*x1 + x2*x2,
// the division and sqrt are IEEE operations (this file is built with -ffp-contract=off: no FMA); log() is the same libm
// function NumPy calls.  tests/test_host_logic.py checks fields, odd counts, cached-gaussian hand-over and the
// continued stream against np.random itself.
#include <immintrin.h>
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

// Instruction-set variants are picked at run time (GCC function multi-versioning).  ROD_RNG_ISA_AVX2 / ROD_RNG_ISA_DEFAULT
// pin one variant at compile time: tests/test_host_logic.py builds those two and checks them against np.random as well, so
// the paths a machine without AVX-512 / AVX2 would take are tested on any machine.
#if defined(ROD_RNG_ISA_DEFAULT)
#define ROD_RNG_CLONES_WIDE
#define ROD_RNG_CLONES
#elif defined(ROD_RNG_ISA_AVX2)
#define ROD_RNG_CLONES_WIDE __attribute__((target("avx2")))
#define ROD_RNG_CLONES __attribute__((target("avx2")))
#else
#define ROD_RNG_CLONES_WIDE __attribute__((target_clones("avx512f", "avx2", "default")))
#define ROD_RNG_CLONES __attribute__((target_clones("avx2", "default")))
#endif

namespace {

constexpr int kN = 624, kM = 397;
constexpr uint32_t kMatrixA = 0x9908b0dfu, kUpper = 0x80000000u, kLower = 0x7fffffffu;

inline uint32_t twist(uint32_t a, uint32_t b) {
    const uint32_t y = (a & kUpper) | (b & kLower);
    return (y >> 1) ^ ((0u - (y & 1u)) & kMatrixA);
}
// nw := the block of 624 raw state words that follows the block `old` (out of place; cloned for AVX2 with run-time
// dispatch: all loops vectorise, the only dependence inside a block has distance 227 words)
ROD_RNG_CLONES void mt_next_generic(const uint32_t* __restrict old, uint32_t* __restrict nw) {
    for (int i = 0; i < kN - kM; ++i) nw[i] = old[i + kM] ^ twist(old[i], old[i + 1]);
    for (int i = kN - kM; i < kN - 1; ++i) nw[i] = nw[i - (kN - kM)] ^ twist(old[i], old[i + 1]);
    nw[kN - 1] = nw[kM - 1] ^ twist(old[kN - 1], nw[0]);
}
// AVX-512: 16 words per step -- the upper / lower bit merge is one vpternlogd, the conditional xor with the matrix constant a
// masked xor (2.3x the AVX2 clone: 1.1 instead of 2.5 ms for the 13 000 blocks of a 1360x765 field).  The last vector of
// either loop overlaps its predecessor (out of place: recomputing a word gives the same word).
__attribute__((target("avx512f"))) inline __m512i mt_step16(const uint32_t* o, __m512i m) {
    const __m512i a = _mm512_loadu_si512(o), b = _mm512_loadu_si512(o + 1);
    const __m512i y = _mm512_ternarylogic_epi32(a, b, _mm512_set1_epi32((int)kUpper), 0xE4);   // mask ? a : b
    const __mmask16 odd = _mm512_test_epi32_mask(y, _mm512_set1_epi32(1));
    const __m512i t = _mm512_xor_si512(_mm512_srli_epi32(y, 1), m);
    return _mm512_mask_xor_epi32(t, odd, t, _mm512_set1_epi32((int)kMatrixA));
}
__attribute__((target("avx512f"))) void mt_next_avx512(const uint32_t* old, uint32_t* nw) {
    constexpr int D = kN - kM;  // 227
    int i = 0;
    for (; i + 16 <= D; i += 16) _mm512_storeu_si512(nw + i, mt_step16(old + i, _mm512_loadu_si512(old + i + kM)));
    i = D - 16;
    _mm512_storeu_si512(nw + i, mt_step16(old + i, _mm512_loadu_si512(old + i + kM)));
    for (i = D; i + 16 <= kN - 1; i += 16) _mm512_storeu_si512(nw + i, mt_step16(old + i, _mm512_loadu_si512(nw + i - D)));
    i = kN - 1 - 16;
    _mm512_storeu_si512(nw + i, mt_step16(old + i, _mm512_loadu_si512(nw + i - D)));
    nw[kN - 1] = nw[kM - 1] ^ twist(old[kN - 1], nw[0]);
}
using MtNextFn = void (*)(const uint32_t*, uint32_t*);
MtNextFn pick_mt_next() {
#if defined(ROD_RNG_ISA_DEFAULT) || defined(ROD_RNG_ISA_AVX2)
    return mt_next_generic;
#else
    __builtin_cpu_init();
    return __builtin_cpu_supports("avx512f") ? mt_next_avx512 : static_cast<MtNextFn>(mt_next_generic);
#endif
}
const MtNextFn mt_next = pick_mt_next();
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

// The polar method of one batch of candidates in three passes, so that everything but the libm log() call runs in vector
// registers (same IEEE operations in the same order as the scalar form in candidate() / legacy_gauss: + - * / sqrt and the
// int -> double / double -> float conversions are exact or correctly rounded either way; no contraction):
//   polar_candidates: words -> x1, x2, r2 of every candidate;  polar_accept: keeps the accepted ones, in order;
//   log() per accepted r2 (scalar);  polar_finish: the two float32 outputs of every accepted pair.
constexpr int kBatch = 1024;   // candidates per batch: 4 arrays of 8 KB, resident in L1
ROD_RNG_CLONES_WIDE
void polar_candidates(const uint32_t* __restrict w, int n, double* __restrict x1, double* __restrict x2, double* __restrict r2) {
    for (int i = 0; i < n; ++i) {
        const uint32_t t0 = temper(w[4 * i]), t1 = temper(w[4 * i + 1]), t2 = temper(w[4 * i + 2]), t3 = temper(w[4 * i + 3]);
        const double d1 = ((int32_t)(t0 >> 5) * 67108864.0 + (int32_t)(t1 >> 6)) / 9007199254740992.0;
        const double d2 = ((int32_t)(t2 >> 5) * 67108864.0 + (int32_t)(t3 >> 6)) / 9007199254740992.0;
        const double a = 2.0 * d1 - 1.0, b = 2.0 * d2 - 1.0;
        x1[i] = a;
        x2[i] = b;
        r2[i] = a * a + b * b;
    }
}
inline int polar_accept(int n, double* x1, double* x2, double* r2) {
    int m = 0;
    for (int i = 0; i < n; ++i) {   // (in place: m <= i)
        const double r = r2[i];
        x1[m] = x1[i];
        x2[m] = x2[i];
        r2[m] = r;
        m += !(r >= 1.0 || r == 0.0) ? 1 : 0;
    }
    return m;
}
ROD_RNG_CLONES_WIDE
void polar_finish(const double* __restrict x1, const double* __restrict x2, const double* __restrict r2, const double* __restrict lg,
                  int m, double sigma, float* __restrict dst) {
    for (int i = 0; i < m; ++i) {
        const double f = __builtin_sqrt(-2.0 * lg[i] / r2[i]);
        dst[2 * i] = (float)(0.0 + sigma * (f * x2[i]));      // returned first
        dst[2 * i + 1] = (float)(0.0 + sigma * (f * x1[i]));  // the "cached" second value
    }
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
    // the generator states at the chunk starts and this thread's chunk of words and of accepted pairs.  Kept per calling
    // thread between calls
    static thread_local std::vector<float> tl_pairs;
    static thread_local std::vector<uint32_t> tl_states, tl_local;
    std::vector<uint32_t>& states = tl_states;  // (local reference: the worker lambdas must see THIS thread's buffer)
    constexpr int kBlocksPerChunk = 52;                                  // 52 blocks of 624 words = 8112 candidates of 4 words
    constexpr uint64_t kChunk = (uint64_t)kBlocksPerChunk * kN / 4;      // candidates per work item
    constexpr size_t kLocalWords = (size_t)(kBlocksPerChunk + 1) * kN;   // a chunk that starts inside a block touches 53
    while (done < n) {
        const uint64_t need_pairs = (n - done + 1) / 2;
        // candidates to draw this round: expectation 4/pi per accepted pair, plus slack; bounded per round
        uint64_t cand = (uint64_t)((double)need_pairs * 1.2740) + 64;
        cand = std::min<uint64_t>(cand, 1ull << 24);
        const int start_pos = *pos;
        // the word stream of this round: the rest of the present block (block 0 = key), then `fresh_blocks` whole blocks.
        // Stream word s is word (start_pos + s) % 624 of block (start_pos + s) / 624.
        const uint64_t in_first = (uint64_t)(kN - start_pos);
        const uint64_t fresh_blocks = (4 * cand > in_first) ? (4 * cand - in_first + kN - 1) / kN : 0;
        const uint64_t total_words = in_first + fresh_blocks * kN;
        cand = total_words / 4;
        const uint64_t n_chunks = (cand + kChunk - 1) / kChunk;
        const uint64_t base = (uint64_t)(start_pos / kN);   // 1 when the present block is used up (pos == 624), else 0
        const int off = start_pos - (int)base * kN;         // every chunk starts at this word of block base + 52 c
        try {
            if (tl_pairs.size() < 2 * kChunk) tl_pairs.resize(2 * kChunk);
            if (states.size() < n_chunks * kN) states.resize(n_chunks * kN);
            if (tl_local.size() < kLocalWords) tl_local.resize(kLocalWords);
        } catch (const std::bad_alloc&) {
            return ROD_ERR_OOM;  // nothing consumed yet in this round: the generator state is still the caller's
        }
        // prefix[c] = accepted pairs of the chunks before c, valid once prefix_upto >= c: the thread that finishes chunk c
        // waits for it (chunks are handed out in order: it is there or about to be), publishes prefix[c + 1] and copies
        // its pairs -- still in its cache -- to their place in `out`; the need_pairs-th accepted pair ends the round
        std::vector<uint64_t> prefix((size_t)n_chunks + 1, 0);
        std::atomic<uint64_t> states_ready{0}, next_chunk{0}, prefix_upto{0};

        // the raw words of chunk c (candidates [c * kChunk, hi)) from the state at its start -> local[off ...]
        auto fill_chunk = [&](uint64_t c, uint64_t hi, uint32_t* local) {
            const uint64_t n_words = (uint64_t)off + 4 * (hi - c * kChunk);
            const uint64_t n_blocks = (n_words + kN - 1) / kN;   // <= 53
            memcpy(local, &states[(size_t)c * kN], sizeof(uint32_t) * kN);
            for (uint64_t b = 1; b < n_blocks; ++b) mt_next(local + (b - 1) * kN, local + b * kN);
        };
        // one work item: the accepted pairs of chunk c, in order, as float32 outputs in the thread's buffer, then in `out`
        auto process_chunk = [&](uint64_t c, uint32_t* local, float* dstp) {
            const uint64_t lo = c * kChunk, hi = std::min(cand, lo + kChunk);
            fill_chunk(c, hi, local);
            const uint32_t* w = local + off;
            uint64_t cnt = 0;
            alignas(64) double x1[kBatch], x2[kBatch], r2[kBatch], lg[kBatch];
            for (uint64_t i = 0; i < hi - lo; i += kBatch) {
                const int nb = (int)std::min<uint64_t>(kBatch, hi - lo - i);
                polar_candidates(w + 4 * i, nb, x1, x2, r2);
                const int m = polar_accept(nb, x1, x2, r2);
                for (int q = 0; q < m; ++q) lg[q] = log(r2[q]);
                polar_finish(x1, x2, r2, lg, m, sigma, dstp + 2 * cnt);
                cnt += (uint64_t)m;
            }
            while (prefix_upto.load(std::memory_order_acquire) < c) std::this_thread::yield();
            const uint64_t p0 = prefix[(size_t)c];
            prefix[(size_t)c + 1] = p0 + cnt;
            prefix_upto.store(c + 1, std::memory_order_release);
            if (p0 < need_pairs) {
                const uint64_t np = std::min<uint64_t>(cnt, need_pairs - p0);
                const uint64_t o = done + 2 * p0;
                const uint64_t nf = std::min<uint64_t>(2 * np, n - o);  // an odd count drops the very last second value
                memcpy(out + o, dstp, nf * sizeof(float));
            }
        };
        auto work = [&](uint32_t* local, float* local_pairs) {
            for (;;) {
                const uint64_t c = next_chunk.fetch_add(1, std::memory_order_relaxed);
                if (c >= n_chunks) return;
                while (states_ready.load(std::memory_order_acquire) <= c) std::this_thread::yield();
                process_chunk(c, local, local_pairs);
            }
        };
        auto worker = [&]() {
            std::vector<uint32_t> local;
            std::vector<float> local_pairs;
            try {
                local.resize(kLocalWords);
                local_pairs.resize(2 * kChunk);
            } catch (const std::bad_alloc&) {
                return;  // the other threads (the caller at the latest) do this one's chunks
            }
            work(local.data(), local_pairs.data());
        };
        // the recurrence is sequential: this thread runs it while the others already consume the states it publishes
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
            alignas(64) uint32_t pp[2][kN];
            const uint32_t* cur = key;   // block 0
            uint64_t c = 0;
            for (uint64_t b = 0; c < n_chunks; ++b) {
                if (b == base + (uint64_t)kBlocksPerChunk * c) {   // the state at the start of chunk c
                    memcpy(&states[(size_t)c * kN], cur, sizeof(uint32_t) * kN);
                    states_ready.store(++c, std::memory_order_release);
                    if (c == n_chunks) break;
                }
                mt_next(cur, pp[b & 1]);
                cur = pp[b & 1];
            }
        }
        work(tl_local.data(), tl_pairs.data());  // the producer helps with whatever is left (all of it when threads == 1)
        for (auto& th : pool) th.join();

        const uint64_t accepted = prefix[(size_t)n_chunks];
        const uint64_t take_pairs = std::min(accepted, need_pairs);
        uint64_t consumed_cand = cand;  // all of them when this round did not reach the target
        if (take_pairs == need_pairs) {
            // the candidate that produced the last pair: rescan its chunk
            uint64_t c = 0;
            while (prefix[(size_t)c + 1] < need_pairs) ++c;
            uint64_t left = need_pairs - prefix[(size_t)c];
            fill_chunk(c, std::min(cand, (c + 1) * kChunk), tl_local.data());
            const uint32_t* w = tl_local.data() + off;
            uint64_t i = 0;
            Candidate last = candidate(w);
            for (;; ++i) {
                last = candidate(w + 4 * i);
                if (last.ok && --left == 0) break;
            }
            consumed_cand = c * kChunk + i + 1;
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
            // its raw words are the state: block blk + 1 of the round, reached from the nearest chunk state before it
            const uint64_t cs = std::min(n_chunks - 1, (blk + 1 - base) / kBlocksPerChunk);
            uint32_t* pp = tl_local.data();
            memcpy(pp, &states[(size_t)cs * kN], sizeof(uint32_t) * kN);
            uint64_t at = base + (uint64_t)kBlocksPerChunk * cs, flip = 0;
            for (; at < blk + 1; ++at, flip ^= 1) mt_next(pp + flip * kN, pp + (flip ^ 1) * kN);
            memcpy(key, pp + flip * kN, sizeof(uint32_t) * kN);
            *pos = (int)(beyond - blk * kN);
        }
    }
    return ROD_OK;
}
