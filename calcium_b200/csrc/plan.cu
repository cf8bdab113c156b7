// plan.cu -- host-side slot planner for K1 (no device code).
//
// K1 gathers, for every stored Laplacian entry, a 16-byte (c, p) pair from shared memory with one
// LDS.128 per warp.  The hardware serves such a load in four phases of eight lanes; inside a phase
// two lanes conflict when they address different 16-byte entries that live in the same group of
// four banks (entry index mod 8).  With cells dealt to threads in arbitrary order a phase costs
// ~2 wavefronts on average and shared-memory bandwidth, not the FP64 pipe, bounds the kernel
// (measured: profiles/r1_k1_baseline.md).
//
// Two static degrees of freedom remove most of that, neither of which touches the arithmetic:
//   1. the cell -> slot map (which thread owns which cell);
//   2. with COPIES == 2 the kernel keeps every value pair twice, in bank groups 4 apart, and each
//      (cell, entry) reads whichever copy collides less ("power of two choices").
// The geometry is shared by every pouch and all 18 000 steps, so the plan is computed once on the
// host: start from the length-sorted order (warps stay uniform in row length: cells only ever move
// within their length class) and anneal over swaps, minimising the number of wavefronts
//     sum over phase groups g (8 consecutive slots) and entry positions e of
//         max over bank groups b of #{distinct entries in bank group b addressed by g at e},
// where the copy choice inside each (g, e) is made greedily with one refinement pass.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

#include "common.cuh"

namespace cab200 {

void default_shape(int n, int *threads, int *cpt);   // integrate.cu

namespace {

struct Planner {
    int n, np, copies, cpt;
    const int32_t *rowptr, *col;
    std::vector<int> slot_of, cell_of;        // cell_of[slot] = cell or -1
    std::vector<std::vector<int>> refs;       // refs[c] = cells whose row contains c (incl. c)
    std::vector<int> wlen;                    // per warp (32 slots): longest row
    int n_groups;

    // Wavefronts of phase group g at entry position e; if sel_out, also reports which lanes should
    // read copy B (bit l of *sel_out).
    inline int phase_cost(int g, int e, unsigned *sel_out) const
    {
        int a_idx[8], b_idx[8];     // candidate entry indices (b = -1: no choice)
        for (int l = 0; l < 8; ++l) {
            const int c = cell_of[g * 8 + l];
            a_idx[l] = (copies + 1) * np;   // zero entry
            b_idx[l] = -1;
            if (c < 0) continue;
            const int rb = rowptr[c], len = rowptr[c + 1] - rb;
            if (e >= len) continue;
            const int cc = col[rb + e];
            if (cc == c) { a_idx[l] = copies * np + slot_of[c]; continue; }
            a_idx[l] = slot_of[cc];
            if (copies == 2) b_idx[l] = np + (slot_of[cc] ^ 4);
        }
        int chosen[8];
        int load[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        auto present = [&](int upto, int idx, int skip) {
            for (int j = 0; j < upto; ++j) if (j != skip && chosen[j] == idx) return true;
            return false;
        };
        // forced lanes first, then the lanes with a choice (greedy: lighter bank group)
        for (int l = 0; l < 8; ++l) chosen[l] = -2;
        for (int l = 0; l < 8; ++l)
            if (b_idx[l] < 0) { if (!present(8, a_idx[l], l)) ++load[a_idx[l] & 7]; chosen[l] = a_idx[l]; }
        for (int l = 0; l < 8; ++l) {
            if (b_idx[l] < 0) continue;
            int pick;
            if (present(8, a_idx[l], l)) pick = a_idx[l];
            else if (present(8, b_idx[l], l)) pick = b_idx[l];
            else pick = (load[a_idx[l] & 7] <= load[b_idx[l] & 7]) ? a_idx[l] : b_idx[l];
            if (!present(8, pick, l)) ++load[pick & 7];
            chosen[l] = pick;
        }
        // one refinement pass: move a lane off the heaviest bank group if that helps
        for (int l = 0; l < 8; ++l) {
            if (b_idx[l] < 0) continue;
            const int cur = chosen[l], alt = (cur == a_idx[l]) ? b_idx[l] : a_idx[l];
            if (present(8, cur, l)) continue;                 // shared address: moving gains nothing
            const int lc = load[cur & 7], la = load[alt & 7] + (present(8, alt, l) ? 0 : 1);
            if (la < lc) {
                --load[cur & 7];
                if (!present(8, alt, l)) ++load[alt & 7];
                chosen[l] = alt;
            }
        }
        int mx = 1;
        for (int b = 0; b < 8; ++b) mx = std::max(mx, load[b]);
        if (sel_out) {
            unsigned m = 0;
            for (int l = 0; l < 8; ++l) if (b_idx[l] >= 0 && chosen[l] == b_idx[l]) m |= 1u << l;
            *sel_out = m;
        }
        return mx;
    }

    int group_cost(int g) const
    {
        const int wl = wlen[(g * 8) / (32 * cpt)];
        int cost = 0;
        for (int e = 0; e < wl; ++e) cost += phase_cost(g, e, nullptr);
        return cost;
    }

    long long total_cost() const
    {
        long long c = 0;
        for (int g = 0; g < n_groups; ++g) c += group_cost(g);
        return c;
    }

    // a warp owns 32*cpt consecutive slots and runs one trip count for all of them: its longest
    // row, rounded up to even (the kernel walks row positions in pairs)
    void compute_wlen()
    {
        const int per = 32 * cpt, nw = (np + per - 1) / per;
        wlen.assign(nw, 0);
        for (int w = 0; w < nw; ++w) {
            for (int s = w * per; s < std::min(np, (w + 1) * per); ++s)
                if (cell_of[s] >= 0) wlen[w] = std::max(wlen[w], rowptr[cell_of[s] + 1] - rowptr[cell_of[s]]);
            wlen[w] = (wlen[w] + 1) & ~1;
        }
    }
};

struct Rng {
    uint64_t s;
    uint64_t next()
    {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D2049BB133111Bull;
        return z ^ (z >> 31);
    }
    double unit() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
};

}  // namespace
}  // namespace cab200

using namespace cab200;

// See include/cab200.h.
extern "C" int cab200_plan_slots(int n, const int32_t *rowptr, const int32_t *colidx, int copies,
                                 int cells_per_thread, long long iters, unsigned long long seed,
                                 int use_input, uint16_t *plan_out, int64_t *cost_out)
{
    if (n <= 0 || n > 32767 || !rowptr || !colidx || !plan_out || (copies != 1 && copies != 2) ||
        cells_per_thread < 1 || cells_per_thread > 8) {
        set_error("cab200_plan_slots: bad argument");
        return CAB200_EINVAL;
    }
    for (int i = 0; i < n; ++i) {
        if (rowptr[i + 1] < rowptr[i] || rowptr[i + 1] - rowptr[i] > 16) {
            set_error("cab200_plan_slots: row %d: length outside [0,16]", i);
            return CAB200_EINVAL;
        }
        for (int j = rowptr[i]; j < rowptr[i + 1]; ++j)
            if (colidx[j] < 0 || colidx[j] >= n) { set_error("cab200_plan_slots: column out of range"); return CAB200_EINVAL; }
    }
    Planner P;
    P.n = n; P.np = (n + 31) & ~31; P.copies = copies; P.cpt = cells_per_thread;
    P.rowptr = rowptr; P.col = colidx;
    P.n_groups = P.np / 8;
    P.slot_of.assign(n, 0);
    P.cell_of.assign(P.np, -1);
    // start: stable sort by row length, descending (what the kernel does on its own)
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        return (rowptr[a + 1] - rowptr[a]) > (rowptr[b + 1] - rowptr[b]);
    });
    if (use_input) {
        std::vector<char> seen(n, 0);
        for (int i = 0; i < n; ++i) {
            const int sl = plan_out[i];
            if (sl >= n || seen[sl]) { set_error("cab200_plan_slots: input is not a permutation"); return CAB200_EINVAL; }
            seen[sl] = 1;
            order[sl] = i;
        }
    }
    for (int s = 0; s < n; ++s) { P.slot_of[order[s]] = s; P.cell_of[s] = order[s]; }
    P.compute_wlen();
    P.refs.assign(n, {});
    for (int i = 0; i < n; ++i)
        for (int j = rowptr[i]; j < rowptr[i + 1]; ++j) P.refs[colidx[j]].push_back(i);

    // length classes: a swap never changes which row lengths a warp holds
    std::vector<std::vector<int>> class_slots(32);
    std::vector<int> class_of(n);
    for (int s = 0; s < n; ++s) {
        const int len = rowptr[order[s] + 1] - rowptr[order[s]];
        class_of[s] = len;
        class_slots[len].push_back(s);
    }

    long long cost = P.total_cost();
    const long long cost0 = cost;
    Rng rng{seed ? seed : 0x5EEDull};
    std::vector<int> touched;
    const double t_start = 0.8, t_end = 0.02;
    for (long long it = 0; it < iters; ++it) {
        const int sa = (int)(rng.next() % (uint64_t)n);
        const std::vector<int> &cls = class_slots[class_of[sa]];
        if (cls.size() < 2) continue;
        const int sb = cls[rng.next() % cls.size()];
        if (sb == sa) continue;
        const int a = P.cell_of[sa], b = P.cell_of[sb];
        touched.clear();
        for (int r : P.refs[a]) touched.push_back(P.slot_of[r] / 8);
        for (int r : P.refs[b]) touched.push_back(P.slot_of[r] / 8);
        touched.push_back(sa / 8); touched.push_back(sb / 8);
        std::sort(touched.begin(), touched.end());
        touched.erase(std::unique(touched.begin(), touched.end()), touched.end());
        int old_c = 0, new_c = 0;
        for (int g : touched) old_c += P.group_cost(g);
        std::swap(P.slot_of[a], P.slot_of[b]);
        P.cell_of[sa] = b; P.cell_of[sb] = a;
        for (int g : touched) new_c += P.group_cost(g);
        const int delta = new_c - old_c;
        const double temp = t_start * std::pow(t_end / t_start, (double)it / (double)iters);
        if (delta <= 0 || rng.unit() < std::exp(-delta / temp)) {
            cost += delta;
        } else {
            std::swap(P.slot_of[a], P.slot_of[b]);
            P.cell_of[sa] = a; P.cell_of[sb] = b;
        }
    }
    // emit: slot_of[n], then per-cell copy-selection bit masks (bit e: entry e reads copy B)
    for (int i = 0; i < n; ++i) { plan_out[i] = (uint16_t)P.slot_of[i]; plan_out[n + i] = 0; }
    if (copies == 2) {
        for (int g = 0; g < P.n_groups; ++g) {
            const int wl = P.wlen[(g * 8) / (32 * P.cpt)];
            for (int e = 0; e < wl; ++e) {
                unsigned m = 0;
                P.phase_cost(g, e, &m);
                for (int l = 0; l < 8; ++l)
                    if ((m >> l) & 1u) plan_out[n + P.cell_of[g * 8 + l]] |= (uint16_t)(1u << e);
            }
        }
    }
    if (cost_out) {
        cost_out[0] = cost0; cost_out[1] = cost; cost_out[2] = P.total_cost();
        long long rows = 0;      // LDS.128 instructions per step: per warp, trip count x groups it owns
        for (int g = 0; g < P.n_groups; g += 4) rows += P.wlen[(g * 8) / (32 * P.cpt)];
        cost_out[3] = rows;
    }
    return CAB200_OK;
}
