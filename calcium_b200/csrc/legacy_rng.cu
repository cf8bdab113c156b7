// legacy_rng.cu -- host-only helper: numpy's legacy `RandomState.choice(n, k, replace=False)`.
//
// The reference samples its negative prompts with `np.random.choice(len(bg_coords[0]), num_samples,
// replace=False)` (utils/sam2_data_generator.py:163), which numpy's legacy generator implements as
// `permutation(n)[:k]`: a Fisher-Yates shuffle of arange(n) from the top (i = n-1 .. 1, j = random_interval(i),
// swap), every j drawn with the masked-rejection `random_interval` on 32-bit MT19937 outputs.  For the ~10^5
// candidate pixels of a frame that is 10^5 generator calls per frame: numpy needs 2-5 ms for it and holds the
// interpreter lock, which makes it the largest single cost of the batch path once everything else runs on the
// GPU.  This is the same algorithm on the same generator state, word for word (numpy/random/src/mt19937,
// legacy-distributions `random_interval`, `_shuffle_raw` of mtrand.pyx), as one tight loop outside the
// interpreter lock: same indices out, same generator state afterwards (tests/test_prompts.py compares both with
// numpy itself).
#include <stdint.h>
#include <stdlib.h>
#if defined(__GNUC__) && defined(__x86_64__)
#include <immintrin.h>
#endif

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace {

constexpr int MT_N = 624, MT_M = 397;
constexpr uint32_t MATRIX_A = 0x9908b0dfu, UPPER_MASK = 0x80000000u, LOWER_MASK = 0x7fffffffu;

// The state update and the output tempering as plain loops the host compiler vectorises; compiled twice -- for
// the baseline x86-64 target and with AVX2 (eight outputs per instruction) -- and picked once at run time.  Same
// integer arithmetic either way.
#if defined(__GNUC__) && defined(__x86_64__)
#define CAB_TARGET_AVX2 __attribute__((target("avx2")))
#define CAB_ALWAYS_INLINE inline __attribute__((always_inline))
#else
#define CAB_TARGET_AVX2
#define CAB_ALWAYS_INLINE inline
#endif

static CAB_ALWAYS_INLINE void mt_gen_body(uint32_t *__restrict key)
{
    int i;
    // key[i] depends on key[i + 1] and key[i + 397] (not yet updated) or key[i - 227] (updated 227 iterations ago)
    for (i = 0; i < MT_N - MT_M; i++) {
        const uint32_t y = (key[i] & UPPER_MASK) | (key[i + 1] & LOWER_MASK);
        key[i] = key[i + MT_M] ^ (y >> 1) ^ (-(y & 1) & MATRIX_A);
    }
    for (; i < MT_N - 1; i++) {
        const uint32_t y = (key[i] & UPPER_MASK) | (key[i + 1] & LOWER_MASK);
        key[i] = key[i + (MT_M - MT_N)] ^ (y >> 1) ^ (-(y & 1) & MATRIX_A);
    }
    const uint32_t y = (key[MT_N - 1] & UPPER_MASK) | (key[0] & LOWER_MASK);
    key[MT_N - 1] = key[MT_M - 1] ^ (y >> 1) ^ (-(y & 1) & MATRIX_A);
}
static CAB_ALWAYS_INLINE void mt_temper_body(const uint32_t *__restrict key, uint32_t *__restrict out, int from)
{
    for (int i = from; i < MT_N; ++i) {
        uint32_t y = key[i];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        out[i] = y;
    }
}
static void mt_gen_generic(uint32_t *key) { mt_gen_body(key); }
static void mt_temper_generic(const uint32_t *key, uint32_t *out, int from) { mt_temper_body(key, out, from); }
CAB_TARGET_AVX2 static void mt_gen_avx2(uint32_t *key) { mt_gen_body(key); }
CAB_TARGET_AVX2 static void mt_temper_avx2(const uint32_t *key, uint32_t *out, int from) { mt_temper_body(key, out, from); }
static bool mt_use_avx2()
{
#if defined(__GNUC__) && defined(__x86_64__)
    static const bool yes = __builtin_cpu_supports("avx2");
    return yes;
#else
    return false;
#endif
}

struct Mt {
    uint32_t *key;
    int pos;
    void gen()
    {
        if (mt_use_avx2()) mt_gen_avx2(key); else mt_gen_generic(key);
        pos = 0;
    }
    inline uint32_t next32()
    {
        if (pos == MT_N) gen();
        uint32_t y = key[pos++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
};

}  // namespace

namespace {

// Tempered outputs of the rest of the current state block, produced a block at a time (a plain loop the host
// compiler vectorises) instead of one output per call; `mt.pos` is advanced as outputs are consumed, so the
// generator state left behind is exactly what per-call draws would leave.
struct MtBlock {
    Mt &mt;
    uint32_t out[MT_N];
    int end = 0;                      // outputs out[mt.pos .. end) are ready
    explicit MtBlock(Mt &m) : mt(m) { fill(); }
    void fill()
    {
        if (mt.pos == MT_N) mt.gen();
        if (mt_use_avx2()) mt_temper_avx2(mt.key, out, mt.pos); else mt_temper_generic(mt.key, out, mt.pos);
        end = MT_N;
    }
    inline uint32_t next32()
    {
        if (mt.pos == end) fill();
        return out[mt.pos++];
    }
};

// 64 masked outputs against the window [sure, top] (see choice_first_k): -1 when one of them falls into the band
// (sure, top], else the accepted ones (<= sure) are written to dst in order and their number is returned.
static int block64_generic(const uint32_t *o, uint32_t mask, uint32_t top, uint32_t sure, uint32_t *dst)
{
    uint32_t v[64];
    uint32_t band = 0;
    for (int j = 0; j < 64; ++j) {
        v[j] = o[j] & mask;
        band |= (uint32_t)(v[j] > sure) & (uint32_t)(v[j] <= top);
    }
    if (band) return -1;
    int cnt = 0;
    for (int j = 0; j < 64; ++j) { dst[cnt] = v[j]; cnt += (v[j] <= sure) ? 1 : 0; }      // a rejected one is overwritten
    return cnt;
}
#if defined(__GNUC__) && defined(__x86_64__)
__attribute__((target("avx512f"))) static int block64_avx512(const uint32_t *o, uint32_t mask, uint32_t top, uint32_t sure, uint32_t *dst)
{
    const __m512i m = _mm512_set1_epi32((int)mask), t = _mm512_set1_epi32((int)top), s = _mm512_set1_epi32((int)sure);
    __m512i v[4];
    __mmask16 band = 0;
    for (int q = 0; q < 4; ++q) {
        v[q] = _mm512_and_si512(_mm512_loadu_si512((const void *)(o + 16 * q)), m);
        band |= _mm512_cmpgt_epu32_mask(v[q], s) & _mm512_cmple_epu32_mask(v[q], t);
    }
    if (band) return -1;
    int cnt = 0;
    for (int q = 0; q < 4; ++q) {
        const __mmask16 acc = _mm512_cmple_epu32_mask(v[q], s);
        _mm512_mask_compressstoreu_epi32((void *)(dst + cnt), acc, v[q]);
        cnt += __builtin_popcount((unsigned)acc);
    }
    return cnt;
}
static bool block64_use_avx512()
{
    static const bool yes = __builtin_cpu_supports("avx512f") && !getenv("CAB200_NO_AVX512");     // the variable: tests of the generic path
    return yes;
}
#else
static int block64_avx512(const uint32_t *o, uint32_t mask, uint32_t top, uint32_t sure, uint32_t *dst) { return block64_generic(o, mask, top, sure, dst); }
static bool block64_use_avx512() { return false; }
#endif

// permutation(n)[:k] on the generator mt: numpy's _shuffle_raw + random_interval (see the file header).
// The shuffle runs i = n-1 .. 1, swap(a[i], a[j_i]); only a[0..k) is wanted.  Instead of swapping inside a
// megabyte-sized array (random accesses), the draws j_i are stored in generator order and the k wanted positions
// are traced BACKWARDS through the swaps (i = 1 .. n-1: a position equal to i moves to j_i, one equal to j_i moves
// to i): where a position ends up is the original index = the value numpy returns.  j_i <= i, so beyond i = k a
// tracked position can only be hit as j_i -- k compares per draw over a sequential array.
void choice_first_k(Mt &mt, int64_t n, int32_t k, int64_t *out)
{
    // one scratch array per host thread, kept between calls: a fresh 1 MB vector per frame means an mmap / munmap
    // pair per call, and with 32 prompt threads those serialise on the process's address-space lock.  The draws are
    // kept in GENERATOR order: Jr[t] is the draw for i = n - 1 - t.
    thread_local std::vector<uint32_t> Jr;
    Jr.resize((size_t)n + 64);
    MtBlock blk(mt);
    // random_interval(i): smallest bit mask >= i, rejection on masked 32-bit outputs (i <= 0x7fffffff here).  The
    // accept / reject outcome is close to a coin flip for i just above a power of two, so nothing here branches on
    // it.  Outputs are taken 64 at a time: output j of a block is compared with i minus the number of outputs
    // accepted before it, which lies in [i - 63, i] -- so an output <= i - 63 is accepted and one > i rejected
    // whatever came before, and unless some output falls into the narrow band between (probability 64 / mask each)
    // the block reduces to a compaction of the accepted outputs (AVX-512: four compress-stores).
    {
        int64_t i = n - 1;
        uint32_t *Jp = Jr.data();
        const bool wide = block64_use_avx512();
        while (i >= 1) {
            uint32_t mask = (uint32_t)i;
            mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
            const int64_t lo = std::max<int64_t>(1, (int64_t)(mask >> 1) + 1);     // the mask serves i = lo .. mask
            while (i >= lo) {
                if (mt.pos == blk.end) blk.fill();
                constexpr int B = 64;
                if (blk.end - mt.pos >= B && i - lo + 1 >= B) {
                    const uint32_t top = (uint32_t)i, sure = (uint32_t)(i - (B - 1));
                    uint32_t *dst = Jp + (n - 1 - i);
                    const int cnt = wide ? block64_avx512(blk.out + mt.pos, mask, top, sure, dst)
                                         : block64_generic(blk.out + mt.pos, mask, top, sure, dst);
                    if (cnt >= 0) {
                        mt.pos += B;
                        i -= cnt;
                        continue;
                    }
                }
                int p = mt.pos;
                const int e = std::min(blk.end, p + B);
                while (p < e && i >= lo) {
                    const uint32_t w = blk.out[p++] & mask;
                    Jp[n - 1 - i] = w;
                    i -= (w <= (uint32_t)i) ? 1 : 0;
                }
                mt.pos = p;
            }
        }
    }
    auto draw = [&](int64_t i) -> uint32_t { return Jr[(size_t)(n - 1 - i)]; };      // the draw j_i of swap(a[i], a[j_i])
    constexpr int KMAX = 32;
    if (k <= KMAX) {
        uint32_t pos[KMAX];
        for (int32_t q = 0; q < k; ++q) pos[q] = (uint32_t)q;
        const int64_t head = std::min<int64_t>(n, (int64_t)k);
        for (int64_t i = 1; i < head; ++i) {                 // i < k: a tracked position can also BE i
            const uint32_t j = draw(i);
            for (int32_t q = 0; q < k; ++q) {
                if (pos[q] == (uint32_t)i) pos[q] = j;
                else if (pos[q] == j) pos[q] = (uint32_t)i;
            }
        }
        if (k == 5) {                                        // the reference's num_samples
            uint32_t p0 = pos[0], p1 = pos[1], p2 = pos[2], p3 = pos[3], p4 = pos[4];
            // a hit is rare (k / i per draw): test 16 draws at a time (a reduction the compiler vectorises) and
            // walk a chunk one by one only when it holds a hit
            const uint32_t *Jp = Jr.data();
            int64_t i = head;
            for (; i + 16 <= n; i += 16) {
                const uint32_t *c = Jp + (n - 1 - (i + 15));       // draws of i + 15 .. i (generator order)
                uint32_t hit = 0;
                for (int e = 0; e < 16; ++e) {
                    const uint32_t j = c[e];
                    hit |= (uint32_t)(j == p0) | (uint32_t)(j == p1) | (uint32_t)(j == p2) | (uint32_t)(j == p3) | (uint32_t)(j == p4);
                }
                if (hit) {
                    for (int e = 0; e < 16; ++e) {
                        const uint32_t j = c[15 - e], ii = (uint32_t)(i + e);
                        p0 = (j == p0) ? ii : p0; p1 = (j == p1) ? ii : p1; p2 = (j == p2) ? ii : p2;
                        p3 = (j == p3) ? ii : p3; p4 = (j == p4) ? ii : p4;
                    }
                }
            }
            for (; i < n; ++i) {
                const uint32_t j = draw(i), ii = (uint32_t)i;
                p0 = (j == p0) ? ii : p0; p1 = (j == p1) ? ii : p1; p2 = (j == p2) ? ii : p2;
                p3 = (j == p3) ? ii : p3; p4 = (j == p4) ? ii : p4;
            }
            pos[0] = p0; pos[1] = p1; pos[2] = p2; pos[3] = p3; pos[4] = p4;
        } else {
            for (int64_t i = head; i < n; ++i) {
                const uint32_t j = draw(i);
                for (int32_t q = 0; q < k; ++q) pos[q] = (j == pos[q]) ? (uint32_t)i : pos[q];
            }
        }
        for (int32_t q = 0; q < k; ++q) out[q] = (int64_t)pos[q];
        return;
    }
    // many samples: the shuffle itself, replayed from the stored draws
    thread_local std::vector<int32_t> a;
    a.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) a[(size_t)i] = (int32_t)i;
    for (int64_t i = n - 1; i >= 1; --i) std::swap(a[(size_t)i], a[draw(i)]);
    for (int32_t q = 0; q < k; ++q) out[q] = a[(size_t)q];
}

}  // namespace

namespace cab200 {
// for prompts_host.cu: the same draws on a caller-held state (key[624], pos)
void legacy_choice_first_k(uint32_t *mt_key, int32_t *mt_pos, int64_t n, int32_t k, int64_t *out)
{
    Mt mt{mt_key, *mt_pos};
    choice_first_k(mt, n, k, out);
    *mt_pos = mt.pos;
}
uint32_t legacy_next32(uint32_t *mt_key, int32_t *mt_pos)
{
    Mt mt{mt_key, *mt_pos};
    const uint32_t v = mt.next32();
    *mt_pos = mt.pos;
    return v;
}
}  // namespace cab200

// See include/cab200.h.
extern "C" int cab200_legacy_choice(uint32_t *mt_key, int32_t *mt_pos, int64_t n, int32_t k, int64_t *out)
{
    using cab200::set_error;
    if (!mt_key || !mt_pos || !out || n < 1 || k < 0 || k > n || n > 0x7fffffffLL || *mt_pos < 0 || *mt_pos > MT_N) {
        set_error("cab200_legacy_choice: bad argument");
        return CAB200_EINVAL;
    }
    Mt mt{mt_key, *mt_pos};
    choice_first_k(mt, n, k, out);
    *mt_pos = mt.pos;
    return CAB200_OK;
}

// See include/cab200.h.
extern "C" int cab200_sample_background(const uint16_t *labels, const uint8_t *labelled_lut, int32_t n_lut,
                                        const int32_t *d2, int64_t npix, int32_t min_d2, uint32_t *mt_key,
                                        int32_t *mt_pos, int32_t k, int64_t *px_out, int64_t *count_out)
{
    using cab200::set_error;
    if (!labels || !labelled_lut || !d2 || !mt_key || !mt_pos || !px_out || !count_out || npix < 1 ||
        npix > 0x7fffffffLL || k < 0 || n_lut < 1 || *mt_pos < 0 || *mt_pos > MT_N) {
        set_error("cab200_sample_background: bad argument");
        return CAB200_EINVAL;
    }
    // candidates, row-major: pixels of no instance at least sqrt(min_d2) away from every instance (np.where order)
    int64_t count = 0;
    for (int64_t p = 0; p < npix; ++p) {
        const uint32_t l = labels[p];
        if (l >= (uint32_t)n_lut) { set_error("cab200_sample_background: label outside the table"); return CAB200_EINVAL; }
        count += (!labelled_lut[l] && d2[p] >= min_d2) ? 1 : 0;
    }
    *count_out = count;
    if (count == 0 || k == 0) return CAB200_OK;
    const int32_t kk = (int32_t)std::min<int64_t>(k, count);
    std::vector<int64_t> rank((size_t)kk);
    Mt mt{mt_key, *mt_pos};
    choice_first_k(mt, count, kk, rank.data());
    *mt_pos = mt.pos;
    // rank -> pixel: one more pass, visiting the wanted ranks in ascending order
    std::vector<int32_t> order((size_t)kk);
    for (int32_t q = 0; q < kk; ++q) order[(size_t)q] = q;
    std::sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return rank[(size_t)x] < rank[(size_t)y]; });
    int64_t seen = 0;
    int32_t next = 0;
    for (int64_t p = 0; p < npix && next < kk; ++p) {
        if (!labelled_lut[labels[p]] && d2[p] >= min_d2) {
            while (next < kk && rank[(size_t)order[(size_t)next]] == seen) px_out[order[(size_t)next++]] = p;
            ++seen;
        }
    }
    return CAB200_OK;
}
