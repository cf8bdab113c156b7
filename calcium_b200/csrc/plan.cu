// plan.cu -- host-side layout planner for K1 (no device code).
//
// K1 gathers, for every stored Laplacian entry, a 16-byte (c, p) pair from shared memory with one
// LDS.128 per warp.  The hardware serves such a load in four phases of eight lanes; inside a phase
// two lanes conflict when they address different 16-byte entries that live in the same group of
// four banks (entry index mod 8).  With cells dealt to threads in arbitrary order a phase costs
// ~2 wavefronts on average and shared-memory bandwidth bounds the kernel
// (measured: profiles/r1_k1_v0_baseline_ncu.txt).
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
//
// For pouches too large for one CTA (or to cut single-pouch latency) the cells are first split into
// NCTA contiguous parts along a breadth-first order of the graph (thin boundaries), one part per
// CTA of a thread-block cluster; a row entry whose column lives in another part reads a *halo*
// entry that the owner CTA pushes through distributed shared memory every step.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <numeric>
#include <queue>
#include <vector>

#include "common.cuh"
#include "k1_layout.cuh"

namespace cab200 {

namespace {

// One CTA's planning problem: nloc local cells; row entries are local cell ids (>= 0), or -1-h for
// halo entry h.  The cell's own id marks the diagonal.  Rows are laid out among the 2*ep gather
// positions by k1_row (k1_layout.cuh) -- the same rule the kernel applies.
struct Problem {
    int n, np, copies, cpt, ep;
    int pl = 8;                               // lanes served together by one shared-memory phase = number of bank groups an
                                              // entry can fall into: 8 for 16-byte entries (FP64, LDS.128), 16 for 8-byte ones (FP32, LDS.64)
    bool unit;
    std::vector<int> rowptr, ecol;
    std::vector<K1Row> row;                   // laid out for the diagonal point of the warp the cell sits in
    std::vector<int> rdiag;                   // per cell: index of the diagonal inside the row, -1 none, -2 several
    std::vector<int> slot_of, cell_of;        // cell_of[slot] = cell or -1
    std::vector<std::vector<int>> refs;       // refs[c] = local cells whose row contains c (incl. c)
    std::vector<int> wstart, wend;            // per warp (32*cpt slots): the position range it runs
    std::vector<int> wdiag;                   // per warp: its diagonal point D (k1_warp_diag of its cells)
    int n_groups;
    bool feasible = true;                     // every warp has a diagonal point (set_order)

    inline int rowlen(int c) const { return rowptr[c + 1] - rowptr[c]; }
    inline int nb(int c) const { return k1_row_nb(rowlen(c), rdiag[c]); }
    inline int na(int c) const { return k1_row_na(rowlen(c), rdiag[c]); }
    // a row that does not fit the default diagonal point: it pins its warp's D, and stays in that warp
    inline bool special(int c) const { return unit && (nb(c) > ep || na(c) > ep); }

    void compute_rows()
    {
        row.resize(n);
        rdiag.resize(n);
        for (int c = 0; c < n; ++c) {
            int nd = 0, di = -1;
            for (int j = rowptr[c]; j < rowptr[c + 1]; ++j)
                if (ecol[j] == c) { ++nd; di = j - rowptr[c]; }
            rdiag[c] = nd == 0 ? -1 : (nd == 1 ? di : -2);
            row[c] = k1_row(unit, ep, rowlen(c), rdiag[c], ep);
        }
    }

    // Wavefronts of phase group g at gather position pos; if sel_out, also reports which lanes should
    // read copy B (bit l of *sel_out).
    inline int phase_cost(int g, int pos, unsigned *sel_out) const
    {
        const int zero_idx = copies * np;
        int a_idx[16], b_idx[16];     // candidate entry indices (b = -1: no choice)
        for (int l = 0; l < pl; ++l) {
            const int c = cell_of[g * pl + l];
            a_idx[l] = zero_idx;
            b_idx[l] = -1;
            if (c < 0) continue;
            const int e = k1_entry_at(row[c], ep, pos);
            if (e < 0 || e >= rowlen(c)) continue;
            const int cc = ecol[rowptr[c] + e];
            if (cc < 0) { a_idx[l] = zero_idx + 1 + (-1 - cc); continue; }       // halo entry
            a_idx[l] = slot_of[cc];
            if (copies == 2) b_idx[l] = np + (slot_of[cc] ^ 4);
        }
        int chosen[16];
        int load[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        const int bm = pl - 1;
        auto present = [&](int idx, int skip) {
            for (int j = 0; j < pl; ++j) if (j != skip && chosen[j] == idx) return true;
            return false;
        };
        for (int l = 0; l < pl; ++l) chosen[l] = -2;
        for (int l = 0; l < pl; ++l)      // lanes without a choice first
            if (b_idx[l] < 0) { if (!present(a_idx[l], l)) ++load[a_idx[l] & bm]; chosen[l] = a_idx[l]; }
        for (int l = 0; l < pl; ++l) {    // greedy: the lighter bank group (or an address already read)
            if (b_idx[l] < 0) continue;
            int pick;
            if (present(a_idx[l], l)) pick = a_idx[l];
            else if (present(b_idx[l], l)) pick = b_idx[l];
            else pick = (load[a_idx[l] & bm] <= load[b_idx[l] & bm]) ? a_idx[l] : b_idx[l];
            if (!present(pick, l)) ++load[pick & bm];
            chosen[l] = pick;
        }
        for (int l = 0; l < pl; ++l) {    // one refinement pass
            if (b_idx[l] < 0) continue;
            const int cur = chosen[l], alt = (cur == a_idx[l]) ? b_idx[l] : a_idx[l];
            if (present(cur, l)) continue;
            const int lc = load[cur & bm], la = load[alt & bm] + (present(alt, l) ? 0 : 1);
            if (la < lc) {
                --load[cur & bm];
                if (!present(alt, l)) ++load[alt & bm];
                chosen[l] = alt;
            }
        }
        int mx = 1;
        for (int b = 0; b < pl; ++b) mx = std::max(mx, load[b]);
        if (sel_out) {
            unsigned m = 0;
            for (int l = 0; l < pl; ++l) if (b_idx[l] >= 0 && chosen[l] == b_idx[l]) m |= 1u << l;
            *sel_out = m;
        }
        return mx;
    }

    int group_cost(int g) const
    {
        const int w = (g * pl) / (32 * cpt);
        int cost = 0;
        for (int pos = wstart[w]; pos < wend[w]; ++pos) cost += phase_cost(g, pos, nullptr);
        return cost;
    }

    long long total_cost() const
    {
        long long c = 0;
        for (int g = 0; g < n_groups; ++g) c += group_cost(g);
        return c;
    }

    // a warp owns 32*cpt consecutive slots, has one diagonal point and runs one position range for all of them
    // (k1_warp_diag, k1_warp_range: the rules the kernel applies).  false: some warp's cells do not go together.
    bool compute_wrange()
    {
        const int per = 32 * cpt, nw = (np + per - 1) / per;
        wstart.assign(nw, 0); wend.assign(nw, 0); wdiag.assign(nw, ep);
        bool ok = true;
        for (int w = 0; w < nw; ++w) {
            int mnb = 0, mna = 0;
            for (int s = w * per; s < std::min(np, (w + 1) * per); ++s) {
                const int c = cell_of[s];
                if (c >= 0) { mnb = std::max(mnb, nb(c)); mna = std::max(mna, na(c)); }
            }
            int d = unit ? k1_warp_diag(ep, mnb, mna) : ep;
            if (d < 0) { ok = false; d = ep; }
            wdiag[w] = d;
            int mf = 2 * ep, ml = 0;
            for (int s = w * per; s < std::min(np, (w + 1) * per); ++s) {
                const int c = cell_of[s];
                if (c < 0) continue;
                row[c] = k1_row(unit, ep, rowlen(c), rdiag[c], d);
                if (unit && row[c].legacy) ok = false;            // several diagonal entries: no D fits
                if (rowlen(c) > 0) { mf = std::min(mf, row[c].first); ml = std::max(ml, row[c].last); }
            }
            k1_warp_range(unit, ep, mf, ml, d, &wstart[w], &wend[w]);
        }
        return ok;
    }

    // May cell c (now in warp wc) move to warp w?  Its row, laid out for w's diagonal point, must stay inside w's
    // position range; rows that pin a warp's diagonal point never leave their warp.
    inline bool fits(int c, int wc, int w) const
    {
        if (wc == w) return true;
        if (special(c)) return false;
        if (rowlen(c) == 0) return true;
        const K1Row r = k1_row(unit, ep, rowlen(c), rdiag[c], wdiag[w]);
        return !(unit && r.legacy) && r.first >= wstart[w] && r.last <= wend[w];
    }
    inline void relayout(int c, int w) { if (row[c].dpt != wdiag[w]) row[c] = k1_row(unit, ep, rowlen(c), rdiag[c], wdiag[w]); }

    // order: cell at each slot (a permutation of 0..n-1)
    void set_order(const std::vector<int> &order)
    {
        np = (n + 31) & ~31;
        n_groups = np / pl;
        slot_of.assign(n, 0);
        cell_of.assign(np, -1);
        for (int s = 0; s < n; ++s) { slot_of[order[s]] = s; cell_of[s] = order[s]; }
        feasible = compute_wrange();
        refs.assign(n, {});
        for (int i = 0; i < n; ++i)
            for (int j = rowptr[i]; j < rowptr[i + 1]; ++j)
                if (ecol[j] >= 0) refs[ecol[j]].push_back(i);
    }

    long long gather_instructions() const      // warp-wide loads per step
    {
        long long rows = 0;
        for (int g = 0; g < n_groups; g += 32 / pl) { const int w = (g * pl) / (32 * cpt); rows += wend[w] - wstart[w]; }
        return rows;
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

// Simulated annealing over swaps of two cells that each fit the other's warp (its position range
// covers the row), so no warp's range ever grows.  Returns the final cost.
long long anneal(Problem &P, long long iters, uint64_t seed, long long *cost0_out)
{
    const int n = P.n, per = 32 * P.cpt;
    long long cost = P.total_cost();
    if (cost0_out) *cost0_out = cost;
    Rng rng{seed ? seed : 0x5EEDull};
    std::vector<int> touched;
    // CALCIUM_B200_PLAN_T0: start temperature (default 0.8); a low value polishes an existing plan (use_input)
    const char *t0_env = getenv("CALCIUM_B200_PLAN_T0");
    const double t_start = (t0_env && atof(t0_env) > 0.0) ? atof(t0_env) : 0.8, t_end = 0.02;
    if (n < 2) return cost;
    for (long long it = 0; it < iters; ++it) {
        const int sa = (int)(rng.next() % (uint64_t)n);
        const int sb = (int)(rng.next() % (uint64_t)n);
        if (sb == sa) continue;
        const int a = P.cell_of[sa], b = P.cell_of[sb];
        const int wa = sa / per, wb = sb / per;
        if (wa != wb && !(P.fits(a, wa, wb) && P.fits(b, wb, wa))) continue;
        touched.clear();
        for (int r : P.refs[a]) touched.push_back(P.slot_of[r] / P.pl);
        for (int r : P.refs[b]) touched.push_back(P.slot_of[r] / P.pl);
        touched.push_back(sa / P.pl); touched.push_back(sb / P.pl);
        std::sort(touched.begin(), touched.end());
        touched.erase(std::unique(touched.begin(), touched.end()), touched.end());
        int old_c = 0, new_c = 0;
        for (int g : touched) old_c += P.group_cost(g);
        std::swap(P.slot_of[a], P.slot_of[b]);
        P.cell_of[sa] = b; P.cell_of[sb] = a;
        P.relayout(a, wb); P.relayout(b, wa);
        for (int g : touched) new_c += P.group_cost(g);
        const int delta = new_c - old_c;
        const double temp = t_start * std::pow(t_end / t_start, (double)it / (double)iters);
        if (delta <= 0 || rng.unit() < std::exp(-delta / temp)) {
            cost += delta;
        } else {
            std::swap(P.slot_of[a], P.slot_of[b]);
            P.cell_of[sa] = a; P.cell_of[sb] = b;
            P.relayout(a, wa); P.relayout(b, wb);
        }
    }
    return cost;
}

// copy-selection bit masks (bit e: entry e reads copy B) for the problem's final order
void emit_sel(const Problem &P, std::vector<uint16_t> &sel)
{
    sel.assign(P.n, 0);
    if (P.copies != 2) return;
    for (int g = 0; g < P.n_groups; ++g) {
        const int w = (g * P.pl) / (32 * P.cpt);
        for (int pos = P.wstart[w]; pos < P.wend[w]; ++pos) {
            unsigned m = 0;
            P.phase_cost(g, pos, &m);
            for (int l = 0; l < P.pl; ++l)
                if ((m >> l) & 1u) {
                    const int c = P.cell_of[g * P.pl + l];
                    sel[c] |= (uint16_t)(1u << k1_entry_at(P.row[c], P.ep, pos));     // bit e: row entry e reads copy B
                }
        }
    }
}

int check_csr(int n, const int32_t *rowptr, const int32_t *colidx, const char *who)
{
    for (int i = 0; i < n; ++i) {
        if (rowptr[i + 1] < rowptr[i] || rowptr[i + 1] - rowptr[i] > 16) {
            set_error("%s: row %d: length outside [0,16]", who, i);
            return CAB200_EINVAL;
        }
        for (int j = rowptr[i]; j < rowptr[i + 1]; ++j)
            if (colidx[j] < 0 || colidx[j] >= n) { set_error("%s: column out of range", who); return CAB200_EINVAL; }
    }
    return CAB200_OK;
}

// Start order.  Rows that do not fit the default diagonal point (more than ep entries before, or after, their
// diagonal) go to the END of the slot space: those of one kind share whole warps -- the last kind the last,
// usually partly empty, warp (cells only occupy slots 0 .. n-1) -- topped up with ordinary rows that fit the same
// diagonal point (preferring the ones that do not widen the warp's position range).  The other cells come first,
// with the same (even-rounded) position range next to each other, so that a warp's range is as tight as its cells
// allow.  Four sort keys are tried; the one with the fewest warp-wide loads per step wins.  Returns an empty
// vector when no feasible order was found (e.g. both kinds of special rows in a geometry of a single warp): the
// caller reports it, the geometry then runs the general kernel.
std::vector<int> shape_sorted_order(Problem &P)
{
    const int per = 32 * P.cpt, emax = 2 * P.ep;
    std::vector<char> used(P.n, 0);
    // special rows grouped by the diagonal point they need: the one closest to ep that fits the row itself
    std::vector<std::vector<int>> kinds;
    std::vector<int> kind_d;
    for (int c = 0; c < P.n && P.unit; ++c) {
        if (!P.special(c)) continue;
        const int d = k1_warp_diag(P.ep, P.nb(c), P.na(c));
        if (d < 0) return {};
        size_t ki = 0;
        for (; ki < kind_d.size() && kind_d[ki] != d; ++ki) {}
        if (ki == kind_d.size()) { kind_d.push_back(d); kinds.emplace_back(); }
        kinds[ki].push_back(c);
    }
    const int last_cnt = P.n - ((P.n - 1) / per) * per;       // cells in the last warp that holds any
    std::vector<int> tail;                                    // the sections, in slot order
    int room = P.n;                                           // cells not yet given to a section
    for (size_t ki = kinds.size(); ki-- > 0;) {               // the last kind first: it takes the partial warp
        const std::vector<int> &spec = kinds[ki];
        int mnb = 0, mna = 0;
        for (int c : spec) { mnb = std::max(mnb, P.nb(c)); mna = std::max(mna, P.na(c)); }
        const int D = k1_warp_diag(P.ep, mnb, mna);
        if (D < 0) return {};
        int mf = emax, ml = 0;
        for (int c : spec) {
            const K1Row r = k1_row(true, P.ep, P.rowlen(c), P.rdiag[c], D);
            if (r.legacy) return {};
            mf = std::min(mf, r.first & ~1); ml = std::max(ml, (r.last + 1) & ~1);
            used[c] = 1;
        }
        const bool is_last = ki + 1 == kinds.size();
        int size = is_last ? last_cnt : 0;
        while (size < (int)spec.size()) size += per;
        size = std::min(size, room);
        const int need = size - (int)spec.size();
        if (need < 0) return {};
        std::vector<std::pair<int, int>> cand;                // (range growth in positions, cell)
        for (int c = 0; c < P.n; ++c) {
            if (used[c] || P.special(c) || P.nb(c) > D || P.na(c) > emax - D) continue;
            const K1Row r = k1_row(true, P.ep, P.rowlen(c), P.rdiag[c], D);
            if (r.legacy) continue;
            const int grow = P.rowlen(c) ? std::max(0, mf - (r.first & ~1)) + std::max(0, ((r.last + 1) & ~1) - ml) : 0;
            cand.push_back({grow, c});
        }
        std::stable_sort(cand.begin(), cand.end());
        if ((int)cand.size() < need) return {};
        std::vector<int> sec(spec);
        for (int i = 0; i < need; ++i) { sec.push_back(cand[i].second); used[cand[i].second] = 1; }
        tail.insert(tail.begin(), sec.begin(), sec.end());
        room -= size;
    }
    std::vector<int> best;
    long long best_rows = -1;
    for (int variant = 0; variant < 4; ++variant) {
        std::vector<int> rest;
        for (int c = 0; c < P.n; ++c) if (!used[c]) rest.push_back(c);
        auto row0 = [&](int c) { return k1_row(P.unit, P.ep, P.rowlen(c), P.rdiag[c], P.ep); };
        auto fr = [&](int c) { return P.rowlen(c) ? (row0(c).first & ~1) : 2 * P.ep; };
        auto lr = [&](int c) { return P.rowlen(c) ? ((row0(c).last + 1) & ~1) : 0; };
        std::stable_sort(rest.begin(), rest.end(), [&](int a, int b) {
            const int fa = fr(a), fb = fr(b), la = lr(a), lb = lr(b);
            if (variant == 0) return fa != fb ? fa < fb : la > lb;
            if (variant == 1) return la != lb ? la > lb : fa < fb;
            if (variant == 3) return fa != fb ? fa < fb : la < lb;
            return (la - fa) != (lb - fb) ? (la - fa) > (lb - fb) : fa < fb;
        });
        std::vector<int> order = rest;
        order.insert(order.end(), tail.begin(), tail.end());
        P.set_order(order);
        if (!P.feasible) continue;
        const long long rows = P.gather_instructions();
        if (best_rows < 0 || rows < best_rows) { best_rows = rows; best = order; }
    }
    return best;
}

// number of rows that do not fit the default diagonal point (they pin their warp's); -1 - that number when some
// row has several diagonal entries (no diagonal point fits: general kernel)
int special_rows(int n, const int32_t *rowptr, const int32_t *colidx, int ep)
{
    int cnt = 0;
    bool several = false;
    for (int i = 0; i < n; ++i) {
        int nd = 0, di = -1;
        const int len = rowptr[i + 1] - rowptr[i];
        for (int j = 0; j < len; ++j)
            if (colidx[rowptr[i] + j] == i) { ++nd; di = j; }
        if (nd > 1) { several = true; continue; }
        const int d = nd == 0 ? -1 : di;
        if (k1_row_nb(len, d) > ep || k1_row_na(len, d) > ep) ++cnt;
    }
    return several ? -1 - cnt : cnt;
}

int max_row_of(int n, const int32_t *rowptr)
{
    int m = 0;
    for (int i = 0; i < n; ++i) m = std::max(m, rowptr[i + 1] - rowptr[i]);
    return m;
}

}  // namespace
}  // namespace cab200

using namespace cab200;

// See include/cab200.h.
extern "C" int cab200_plan_slots(int n, const int32_t *rowptr, const int32_t *colidx, int copies, int unit,
                                 int cells_per_thread, long long iters, unsigned long long seed,
                                 int use_input, uint16_t *plan_out, int64_t *cost_out)
{
    return cab200_plan_slots_entry(n, rowptr, colidx, copies, unit, cells_per_thread, 16, iters, seed, use_input, plan_out, cost_out);
}

// See include/cab200.h.
extern "C" int cab200_plan_slots_entry(int n, const int32_t *rowptr, const int32_t *colidx, int copies, int unit,
                                       int cells_per_thread, int entry_bytes, long long iters, unsigned long long seed,
                                       int use_input, uint16_t *plan_out, int64_t *cost_out)
{
    if (n <= 0 || n > 32767 || !rowptr || !colidx || !plan_out || (copies != 1 && copies != 2) ||
        cells_per_thread < 1 || cells_per_thread > 8 || (entry_bytes != 16 && entry_bytes != 8)) {
        set_error("cab200_plan_slots: bad argument");
        return CAB200_EINVAL;
    }
    if (int rc = check_csr(n, rowptr, colidx, "cab200_plan_slots")) return rc;
    Problem P;
    P.n = n; P.copies = copies; P.cpt = cells_per_thread;
    P.pl = 128 / entry_bytes;                 // lanes (and bank groups) of one shared-memory phase
    P.unit = unit != 0; P.ep = k1_ep(max_row_of(n, rowptr));
    P.rowptr.assign(rowptr, rowptr + n + 1);
    P.ecol.assign(colidx, colidx + rowptr[n]);
    P.compute_rows();
    std::vector<int> order = shape_sorted_order(P);
    if (order.empty()) {
        set_error("cab200_plan_slots: no layout with one diagonal point per warp (rows with many entries on either side "
                  "of the diagonal in too small a geometry, or several diagonal entries): run this geometry with unit = 0");
        return CAB200_EINVAL;
    }
    if (use_input) {
        std::vector<char> seen(n, 0);
        for (int i = 0; i < n; ++i) {
            const int sl = plan_out[i];
            if (sl >= n || seen[sl]) { set_error("cab200_plan_slots: input is not a permutation"); return CAB200_EINVAL; }
            seen[sl] = 1;
            order[sl] = i;
        }
    }
    P.set_order(order);
    if (!P.feasible) { set_error("cab200_plan_slots: the given plan puts rows into one warp that share no diagonal point"); return CAB200_EINVAL; }
    long long cost0 = 0;
    const long long cost = anneal(P, iters, seed, &cost0);
    std::vector<uint16_t> sel;
    emit_sel(P, sel);
    for (int i = 0; i < n; ++i) { plan_out[i] = (uint16_t)P.slot_of[i]; plan_out[n + i] = sel[i]; }
    if (cost_out) {
        cost_out[0] = cost0; cost_out[1] = cost; cost_out[2] = P.total_cost();
        cost_out[3] = P.gather_instructions();
    }
    return CAB200_OK;
}

// See include/cab200.h.
extern "C" int cab200_k1_legacy_rows(int n, const int32_t *rowptr, const int32_t *colidx)
{
    if (n <= 0 || !rowptr || !colidx) { set_error("cab200_k1_legacy_rows: bad argument"); return CAB200_EINVAL; }
    return special_rows(n, rowptr, colidx, k1_ep(max_row_of(n, rowptr)));
}

// See include/cab200.h.
extern "C" long long cab200_plan_cluster(int n, const int32_t *rowptr, const int32_t *colidx, int n_cta,
                                         int copies, int cells_per_thread, long long iters,
                                         unsigned long long seed, uint16_t *plan_out,
                                         long long plan_capacity, int64_t *cost_out)
{
    if (n <= 0 || n > 32767 || !rowptr || !colidx || n_cta < 2 || n_cta > 8 || (copies != 1 && copies != 2) ||
        cells_per_thread < 1 || cells_per_thread > 8) {
        set_error("cab200_plan_cluster: bad argument");
        return CAB200_EINVAL;
    }
    if (int rc = check_csr(n, rowptr, colidx, "cab200_plan_cluster")) return rc;
    const int ep = k1_ep(max_row_of(n, rowptr));        // clusters run the unit-weight kernel only
    // ---- 1. breadth-first order from a pseudo-peripheral cell; parts = equal contiguous chunks ----
    auto bfs = [&](int src, std::vector<int> &order) {
        std::vector<char> seen(n, 0);
        order.clear();
        std::queue<int> q;
        auto run = [&](int s0) {
            q.push(s0); seen[s0] = 1;
            while (!q.empty()) {
                const int u = q.front(); q.pop();
                order.push_back(u);
                for (int j = rowptr[u]; j < rowptr[u + 1]; ++j)
                    if (!seen[colidx[j]]) { seen[colidx[j]] = 1; q.push(colidx[j]); }
            }
        };
        run(src);
        for (int i = 0; i < n; ++i) if (!seen[i]) run(i);      // other components, in id order
    };
    std::vector<int> order;
    bfs(0, order);
    bfs(order.back(), order);
    bfs(order.back(), order);
    std::vector<int> part_of(n);
    for (int i = 0; i < n; ++i) part_of[order[i]] = (int)((long long)i * n_cta / n);
    // ---- 2. per part: local ids, halo list, plan ------------------------------------------------
    std::vector<uint16_t> out;
    out.resize(3 * (size_t)n + n_cta);          // part_of | slot_of | sel | halo_count[n_cta] | halo lists
    long long c0 = 0, c1 = 0, rows = 0, halo_total = 0;
    std::vector<std::vector<int>> halos(n_cta);
    for (int q = 0; q < n_cta; ++q) {
        std::vector<int> cells;                 // global ids of this part, ascending
        for (int i = 0; i < n; ++i) if (part_of[i] == q) cells.push_back(i);
        std::vector<int> local_of(n, -1);
        for (size_t k = 0; k < cells.size(); ++k) local_of[cells[k]] = (int)k;
        std::vector<int> &halo = halos[q];
        for (int c : cells)
            for (int j = rowptr[c]; j < rowptr[c + 1]; ++j)
                if (part_of[colidx[j]] != q) halo.push_back(colidx[j]);
        std::sort(halo.begin(), halo.end());
        halo.erase(std::unique(halo.begin(), halo.end()), halo.end());
        Problem P;
        P.n = (int)cells.size(); P.copies = copies; P.cpt = cells_per_thread;
        P.unit = true; P.ep = ep;
        P.rowptr.assign(1, 0);
        for (int c : cells) {
            for (int j = rowptr[c]; j < rowptr[c + 1]; ++j) {
                const int cc = colidx[j];
                if (part_of[cc] == q) P.ecol.push_back(local_of[cc]);
                else P.ecol.push_back(-1 - (int)(std::lower_bound(halo.begin(), halo.end(), cc) - halo.begin()));
            }
            P.rowptr.push_back((int)P.ecol.size());
        }
        if (P.n == 0) { set_error("cab200_plan_cluster: empty part"); return CAB200_EINVAL; }
        P.compute_rows();
        {
            const std::vector<int> order0 = shape_sorted_order(P);
            if (order0.empty()) { set_error("cab200_plan_cluster: part %d has no layout with one diagonal point per warp", q); return CAB200_EINVAL; }
            P.set_order(order0);
        }
        long long cost0 = 0;
        c1 += anneal(P, iters / n_cta, seed + 977 * (q + 1), &cost0);
        c0 += cost0;
        rows += P.gather_instructions();
        std::vector<uint16_t> sel;
        emit_sel(P, sel);
        for (int k = 0; k < P.n; ++k) {
            out[cells[k]] = (uint16_t)q;
            out[n + cells[k]] = (uint16_t)P.slot_of[k];
            out[2 * (size_t)n + cells[k]] = sel[k];
        }
        out[3 * (size_t)n + q] = (uint16_t)halo.size();
        halo_total += (long long)halo.size();
    }
    for (int q = 0; q < n_cta; ++q)
        for (int h : halos[q]) out.push_back((uint16_t)h);
    if (cost_out) { cost_out[0] = c0; cost_out[1] = c1; cost_out[2] = halo_total; cost_out[3] = rows; }
    if (plan_out) {
        if ((long long)out.size() > plan_capacity) {
            set_error("cab200_plan_cluster: plan needs %zu uint16, capacity %lld", out.size(), plan_capacity);
            return CAB200_ESIZE;
        }
        std::copy(out.begin(), out.end(), plan_out);
    }
    return (long long)out.size();
}
