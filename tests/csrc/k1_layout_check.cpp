// TEST INFRASTRUCTURE -- CPU emulation of the gather ORDER of K1 for a given slot plan.
//
// calcium_b200/csrc/k1_layout.cuh holds the row layout rule shared by the kernel (integrate.cuh) and the host
// planner (plan.cu).  This program applies that rule the way the kernel's prologue does -- per warp the diagonal
// point from its cells (k1_warp_diag), per cell the row layout (k1_row), per position the entry (k1_entry_at), the
// warp's position range (k1_warp_range) -- and walks the unrolled position sequence of one step: it checks that every
// stored entry of every row is added exactly once, in ascending column order, with the diagonal term at its place,
// inside the warp's range.  That is the property the kernel's bit-exactness against scipy's csr_matvec rests on.
#include <cstdint>
#include <vector>

#include "../../calcium_b200/csrc/k1_layout.cuh"

using namespace cab200;

// returns 0 if the layout is valid, else 1 + the first offending cell (negative: a structural error code)
extern "C" long long k1_layout_check(int n, const int32_t *rowptr, const int32_t *colidx, const uint16_t *slot_of,
                                     int cells_per_thread, int n_slots, int32_t *warp_diag_out)
{
    int max_row = 0;
    for (int i = 0; i < n; ++i) max_row = rowptr[i + 1] - rowptr[i] > max_row ? rowptr[i + 1] - rowptr[i] : max_row;
    const int EP = k1_ep(max_row), EMAX = 2 * EP, per = 32 * cells_per_thread;
    if (max_row > EMAX) return -1;
    std::vector<int> cell_of(n_slots, -1);
    for (int i = 0; i < n; ++i) {
        if (slot_of[i] >= n_slots || cell_of[slot_of[i]] != -1) return -2;      // not a permutation into the slots
        cell_of[slot_of[i]] = i;
    }
    std::vector<int> rdiag(n), len(n);
    for (int i = 0; i < n; ++i) {
        int nd = 0, di = -1;
        len[i] = rowptr[i + 1] - rowptr[i];
        for (int j = 0; j < len[i]; ++j) {
            if (colidx[rowptr[i] + j] == i) { ++nd; di = j; }
            if (j > 0 && colidx[rowptr[i] + j] <= colidx[rowptr[i] + j - 1]) return -3;   // columns must ascend
        }
        rdiag[i] = nd == 0 ? -1 : (nd == 1 ? di : -2);
    }
    const int n_warps = (n_slots + per - 1) / per;
    for (int w = 0; w < n_warps; ++w) {
        int mnb = 0, mna = 0, any = 0;
        for (int s = w * per; s < (w + 1) * per && s < n_slots; ++s) {
            const int c = cell_of[s];
            if (c < 0) continue;
            any = 1;
            const int nb = k1_row_nb(len[c], rdiag[c]), na = k1_row_na(len[c], rdiag[c]);
            mnb = nb > mnb ? nb : mnb; mna = na > mna ? na : mna;
        }
        const int D = k1_warp_diag(EP, mnb, mna);
        if (warp_diag_out) warp_diag_out[w] = any ? D : 0;
        if (!any) continue;
        if (D < 0) return 1 + cell_of[w * per] ;                                  // this warp's cells share no diagonal point
        int mf = EMAX, ml = 0;
        for (int s = w * per; s < (w + 1) * per && s < n_slots; ++s) {
            const int c = cell_of[s];
            if (c < 0 || len[c] == 0) continue;
            const K1Row r = k1_row(true, EP, len[c], rdiag[c], D);
            if (r.legacy) return 1 + c;
            mf = r.first < mf ? r.first : mf; ml = r.last > ml ? r.last : ml;
        }
        int start, end;
        k1_warp_range(true, EP, mf, ml, D, &start, &end);
        const int dpair = k1_diag_pair(EP, D);
        for (int s = w * per; s < (w + 1) * per && s < n_slots; ++s) {
            const int c = cell_of[s];
            if (c < 0) continue;
            const K1Row r = k1_row(true, EP, len[c], rdiag[c], D);
            // the sequence of additions of one step, as the unrolled pairs make them
            std::vector<int> seq;                                                 // row entry indices; -9 = register diagonal
            for (int p0 = start; p0 < end; p0 += 2) {
                const int e0 = k1_entry_at(r, EP, p0), e1 = k1_entry_at(r, EP, p0 + 1);
                if (e0 >= 0) seq.push_back(e0);
                if ((EP & 1) && dpair == p0 && r.skip) seq.push_back(-9);
                if (e1 >= 0) seq.push_back(e1);
                if (!(EP & 1) && dpair == p0 && r.skip) seq.push_back(-9);
            }
            // every entry outside [start, end) would be lost
            for (int pos = 0; pos < EMAX; ++pos)
                if ((pos < start || pos >= end) && k1_entry_at(r, EP, pos) >= 0) return 1 + c;
            if ((int)seq.size() != len[c]) return 1 + c;
            for (int j = 0; j < len[c]; ++j) {
                const int want = (j == rdiag[c]) ? -9 : j;                        // CSR order, the diagonal at its place
                if (seq[(size_t)j] != want) return 1 + c;
            }
        }
    }
    return 0;
}
