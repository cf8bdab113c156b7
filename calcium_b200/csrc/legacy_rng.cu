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

#include <vector>

#include "common.cuh"

namespace {

constexpr int MT_N = 624, MT_M = 397;
constexpr uint32_t MATRIX_A = 0x9908b0dfu, UPPER_MASK = 0x80000000u, LOWER_MASK = 0x7fffffffu;

struct Mt {
    uint32_t *key;
    int pos;
    void gen()
    {
        int i;
        uint32_t y;
        for (i = 0; i < MT_N - MT_M; i++) {
            y = (key[i] & UPPER_MASK) | (key[i + 1] & LOWER_MASK);
            key[i] = key[i + MT_M] ^ (y >> 1) ^ (-(y & 1) & MATRIX_A);
        }
        for (; i < MT_N - 1; i++) {
            y = (key[i] & UPPER_MASK) | (key[i + 1] & LOWER_MASK);
            key[i] = key[i + (MT_M - MT_N)] ^ (y >> 1) ^ (-(y & 1) & MATRIX_A);
        }
        y = (key[MT_N - 1] & UPPER_MASK) | (key[0] & LOWER_MASK);
        key[MT_N - 1] = key[MT_M - 1] ^ (y >> 1) ^ (-(y & 1) & MATRIX_A);
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

// See include/cab200.h.
extern "C" int cab200_legacy_choice(uint32_t *mt_key, int32_t *mt_pos, int64_t n, int32_t k, int64_t *out)
{
    using cab200::set_error;
    if (!mt_key || !mt_pos || !out || n < 1 || k < 0 || k > n || n > 0x7fffffffLL || *mt_pos < 0 || *mt_pos > MT_N) {
        set_error("cab200_legacy_choice: bad argument");
        return CAB200_EINVAL;
    }
    Mt mt{mt_key, *mt_pos};
    std::vector<int32_t> a((size_t)n);
    for (int64_t i = 0; i < n; ++i) a[(size_t)i] = (int32_t)i;
    for (int64_t i = n - 1; i >= 1; --i) {
        // random_interval(i): smallest bit mask >= i, rejection on masked 32-bit outputs (i <= 0xffffffff here)
        uint32_t mask = (uint32_t)i;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        uint32_t v;
        while ((v = (mt.next32() & mask)) > (uint32_t)i) {}
        const int32_t t = a[v];
        a[v] = a[(size_t)i];
        a[(size_t)i] = t;
    }
    for (int32_t q = 0; q < k; ++q) out[q] = a[(size_t)q];
    *mt_pos = mt.pos;
    return CAB200_OK;
}
