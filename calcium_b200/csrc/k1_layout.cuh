// k1_layout.cuh -- the row layout rule of K1, shared by the kernel (integrate.cuh) and its host
// planner (plan.cu) so that the planner's bank-conflict model sees exactly the accesses the kernel makes.
//
// K1 walks EMAX = 2*EP compile-time gather positions per owned cell.  For the adjacency-minus-degree
// Laplacian (UNIT) the diagonal term L_ii*(c_i, p_i) is NOT gathered from shared memory: the owner has
// c_i and p_i in registers and adds the product at a fixed point of the unrolled sequence, between
// positions EP-1 and EP.  A row's entries before its diagonal are right-aligned below EP, the entries
// after it left-aligned from EP, everything else points at the zero entry (adding +0.0 to a sum that
// started from +0.0 never changes it), so the accumulation order is still the CSR order scipy uses
// (reference core/pouch.py:346-347).  Rows that do not fit -- more than EP entries on one side of the
// diagonal, or several diagonal entries -- keep the round-1 layout ("legacy"): all entries right-aligned
// at EMAX, the diagonal read from the cell's shared-memory diagonal entry, which the owner rewrites every
// step (17 of 1500 cells of the medium disc).  Without UNIT every row is legacy and the diagonal is an
// ordinary weighted entry.
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
    bool legacy;
};

// diag: index inside the row of its only diagonal entry, -1 if it has none, -2 if it has several.
CAB_HD static inline K1Row k1_row(bool unit, int EP, int len, int diag)
{
    K1Row r;
    const int nb = diag >= 0 ? diag : len;
    const int na = diag >= 0 ? len - diag - 1 : 0;
    if (unit && diag != -2 && nb <= EP && na <= EP) {
        r.legacy = false; r.lo_base = EP - nb; r.split = nb; r.skip = diag >= 0 ? 1 : 0;
        r.first = EP - nb; r.last = EP + na;
    } else {
        r.legacy = true; r.lo_base = 2 * EP - len; r.split = len; r.skip = 0;
        r.first = 2 * EP - len; r.last = 2 * EP;
    }
    return r;
}

// Row entry gathered at position pos, or -1 (zero entry).
CAB_HD static inline int k1_entry_at(const K1Row &r, int EP, int pos)
{
    if (pos < r.first || pos >= r.last) return -1;
    if (r.legacy || pos < EP) return pos - r.lo_base;
    return r.split + r.skip + (pos - EP);
}

// Positions are processed in pairs (P0, P0+1), P0 even; the register diagonal is added inside the
// pair (EP-1, EP) when EP is odd and after the pair (EP-2, EP-1) when EP is even.
CAB_HD static inline int k1_diag_pair(int EP) { return (EP & 1) ? EP - 1 : EP - 2; }

// The pair range [start, end) a warp runs, from the smallest `first` and largest `last` of its cells
// (min_first = 2*EP, max_last = 0 for a warp without cells).
CAB_HD static inline void k1_warp_range(bool unit, int EP, int min_first, int max_last, int *start, int *end)
{
    int s = (min_first < 0 ? 0 : min_first) & ~1;
    int e = (max_last + 1) & ~1;
    if (e > 2 * EP) e = 2 * EP;
    if (unit) {                       // the pair that carries the diagonal always runs
        const int dp = k1_diag_pair(EP);
        if (s > dp) s = dp;
        if (e < dp + 2) e = dp + 2;
    }
    if (e < s) e = s;
    *start = s; *end = e;
}

}  // namespace cab200
