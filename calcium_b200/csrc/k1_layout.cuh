// k1_layout.cuh -- the row layout rule of K1, shared by the kernel (integrate.cuh) and its host
// planner (plan.cu) so that the planner's bank-conflict model sees exactly the accesses the kernel makes.
//
// K1 walks EMAX = 2*EP compile-time gather positions per owned cell.  For the adjacency-minus-degree
// Laplacian (UNIT) the diagonal term L_ii*(c_i, p_i) is NOT gathered from shared memory: the owner has
// c_i and p_i in registers and adds the product at a fixed point of the unrolled sequence, after D positions
// (the warp's *diagonal point*).  A row's entries before its diagonal are right-aligned below D, the entries
// after it left-aligned from D, everything else points at the zero entry (adding +0.0 to a sum that
// started from +0.0 never changes it), so the accumulation order is still the CSR order scipy uses
// (reference core/pouch.py:346-347).  D is one value per WARP (a warp-uniform jump decides in which pair of
// positions the product is added): EP for almost every warp; the planner collects the few rows with more than
// EP entries on one side of their diagonal (17 of 1500 cells of the medium disc, all with 6 or 7 entries before
// or after it) in one or two warps whose D is EP + 2 or EP - 2, together with ordinary rows that also fit there.
// Both sides derive D from the warp's cells with k1_warp_diag, so no extra plan field is needed.  A row with
// several diagonal entries fits no D: such a geometry runs the general-weights kernel (UNIT = false), where
// every row is right-aligned at EMAX and the diagonal is an ordinary weighted entry ("legacy" layout).
#pragma once

namespace cab200 {

#ifdef __CUDACC__
#define CAB_HD __host__ __device__
#else
#define CAB_HD
#endif

// Kernel variant (EP) for a geometry whose longest CSR row stores max_row entries.
CAB_HD static inline int k1_ep(int max_row) { return (max_row + 1) / 2 <= 5 ? 5 : 8; }

struct K1Row {
    int lo_base;        // entry e < split sits at position lo_base + e
    int split;          // entries before the register diagonal (legacy: all of them)
    int skip;           // 1: entry `split` is the diagonal, added from registers
    int first, last;    // occupied positions [first, last)
    int dpt;            // the diagonal point D the row was laid out for
    bool legacy;
};

// Entries before / after the diagonal.  diag: index inside the row of its only diagonal entry, -1 if it has none
// (all entries count as "before"), -2 if it has several.
CAB_HD static inline int k1_row_nb(int len, int diag) { return diag >= 0 ? diag : len; }
CAB_HD static inline int k1_row_na(int len, int diag) { return diag >= 0 ? len - diag - 1 : 0; }

// The diagonal point of a warp whose rows have at most max_nb entries before and max_na after the diagonal:
// the D closest to EP with D = EP (mod 2) -- the parity fixes whether the product is added inside a pair of
// positions or after one, a compile-time property of the kernel -- max_nb <= D <= 2*EP - max_na, and D >= 1
// (odd EP) or >= 2 (even EP).  -1: no such D (the warp's cells do not go together).
CAB_HD static inline int k1_warp_diag(int EP, int max_nb, int max_na)
{
    const int lo = max_nb > ((EP & 1) ? 1 : 2) ? max_nb : ((EP & 1) ? 1 : 2);
    const int hi = 2 * EP - max_na;
    for (int k = 0; k <= 2 * EP; k += 2) {
        if (EP + k >= lo && EP + k <= hi) return EP + k;
        if (EP - k >= lo && EP - k <= hi) return EP - k;
    }
    return -1;
}

// D: the diagonal point of the warp that owns the row (k1_warp_diag).
CAB_HD static inline K1Row k1_row(bool unit, int EP, int len, int diag, int D)
{
    K1Row r;
    const int nb = k1_row_nb(len, diag), na = k1_row_na(len, diag);
    r.dpt = D;
    if (unit && diag != -2 && nb <= D && na <= 2 * EP - D) {
        r.legacy = false; r.lo_base = D - nb; r.split = nb; r.skip = diag >= 0 ? 1 : 0;
        r.first = D - nb; r.last = D + na;
    } else {
        r.legacy = true; r.lo_base = 2 * EP - len; r.split = len; r.skip = 0;
        r.first = 2 * EP - len; r.last = 2 * EP;
    }
    return r;
}

// Row entry gathered at position pos, or -1 (zero entry).
CAB_HD static inline int k1_entry_at(const K1Row &r, int EP, int pos)
{
    (void)EP;
    if (pos < r.first || pos >= r.last) return -1;
    if (r.legacy || pos < r.dpt) return pos - r.lo_base;
    return r.split + r.skip + (pos - r.dpt);
}

// Positions are processed in pairs (P0, P0+1), P0 even; the register diagonal is added inside the
// pair (D-1, D) when EP (hence D) is odd and after the pair (D-2, D-1) when it is even.
CAB_HD static inline int k1_diag_pair(int EP, int D) { return (EP & 1) ? D - 1 : D - 2; }

// The pair range [start, end) a warp runs, from the smallest `first` and largest `last` of its cells
// (min_first = 2*EP, max_last = 0 for a warp without cells) and its diagonal point D.
CAB_HD static inline void k1_warp_range(bool unit, int EP, int min_first, int max_last, int D, int *start, int *end)
{
    int s = (min_first < 0 ? 0 : min_first) & ~1;
    int e = (max_last + 1) & ~1;
    if (e > 2 * EP) e = 2 * EP;
    if (unit) {                       // the pair that carries the diagonal always runs
        const int dp = k1_diag_pair(EP, D);
        if (s > dp) s = dp;
        if (e < dp + 2) e = dp + 2;
    }
    if (e < s) e = s;
    *start = s; *end = e;
}

}  // namespace cab200
