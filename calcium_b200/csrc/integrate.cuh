// integrate.cuh -- K1, the persistent ensemble integrator (one CTA per pouch).
//
// Replaces reference core/pouch.py:338-366 (time loop), :31-104 (_simulate_time_step) and the two
// scipy CSR products per step (:346-347) for a whole ensemble in one launch.
//
// Data placement for one pouch (n cells), fixed for all T steps:
//   shared   two buffers (step parity) of (COPIES+1)*NP + 1 entries, one entry = the pair (c, p):
//              [0, NP)            value entries   (c_i, p_i) at index slot_i
//              [NP, 2NP)          (COPIES == 2) a second copy of the values at NP + (slot_i ^ 4),
//                                 i.e. in the bank group 4 away: every gather may read either copy,
//                                 which the host planner uses to dodge bank conflicts
//              [COPIES*NP]        the zero entry (+0, +0), never written after the prologue
//            Step t reads buffer t&1 and writes buffer (t+1)&1: one split barrier per step.
//   regs     per owned cell: c, p, s, r, V_PLC, L_ii and the shared addresses of the row's entries
//            (time-invariant), plus the pouch's parameters (uniform registers).
//   HBM      only the captured frames, [F][4][n] in original cell order.
//
// A Laplacian row is then a branch-free gather: sum = +0; for each stored entry in CSR order
// sum += entry -- off-diagonal entries index the neighbour's value entry (1.0*x == x exactly for the
// adjacency-minus-degree Laplacian), the diagonal term L_ii*(c_i, p_i) is computed from the owner's
// registers and added at its place in the order (after D positions of the unrolled sequence, D the
// warp's diagonal point, see k1_layout.cuh: no shared-memory store or load for it), and unused positions
// read the zero entry.  Because the sum starts from +0.0 it can never be -0.0, so adding the +0.0 padding is an
// exact no-op and the result is bit-identical to scipy's csr_matvec followed by the reference's
// `D * (...)`.
//
// Cell -> thread ownership is static.  Cells are ranked by CSR row length (descending, stable) and
// dealt to slots `k*THREADS + tid`, so the 32 cells a warp handles together have (almost) equal
// row length and the padding costs next to nothing.
#pragma once
#include "common.cuh"
#include "k1_layout.cuh"

namespace cab200 {

struct SimArgs {
    const int32_t *pouch_list;  // [grid] pouch ids handled by this launch
    const int64_t *cell_off;    // [P+1]
    const int32_t *rowptr;      // this geometry's CSR
    const int32_t *colidx;
    const double *vals;
    const uint16_t *slot_plan;  // optional plan: cab200_plan_slots (slot_of[n] | copy_sel[n]) or, for clusters,
                                // cab200_plan_cluster (part_of[n] | slot_of[n] | copy_sel[n] | halo counts | halo lists)
    int n;
    int np;                     // n rounded up to 32: entries per region
    const double *params;
    const double *c0, *p0, *r0, *vplc;
    int T, n_frames;
    const int32_t *frame_steps;
    void *frames_out;
    void *frames_tmp;           // optional: frames are written HERE in slot order (fully coalesced), blocks of
                                // [n_frames][4][NCTA*NSLOT] per pouch of this launch; unpermute_kernel then fills
                                // frames_out.  For captures of (nearly) every step, where the scattered
                                // cell-order stores would dominate the step.
    int32_t *status_out;
};

// Shared-memory bytes K1 needs for a pouch padded to `np` slots.  Entries per step-parity buffer:
// `copies` value regions, the zero entry, the halo.
static inline size_t integrate_buf_entries(int nr, int copies, bool unit, int hmax)
{
    (void)unit;
    return (size_t)copies * nr + 1 + hmax;
}
// persistent table behind the two buffers: orig_of[NSLOT] u16
static inline size_t integrate_tail_bytes(int nslot, size_t elem2, bool unit)
{
    (void)elem2; (void)unit;
    return 2 * (size_t)nslot + 64;
}
static inline size_t integrate_buf_stride(int np, int nslot, size_t elem2, int copies, bool unit, int hmax = 0)
{
    const size_t fixed = (integrate_buf_entries(nslot, copies, unit, hmax) * elem2 + 127) & ~(size_t)127;
    if (2 * fixed + integrate_tail_bytes(nslot, elem2, unit) <= 227 * 1024) return fixed;     // == BufGeom::STRIDE
    return (integrate_buf_entries(np, copies, unit, hmax) * elem2 + 127) & ~(size_t)127;       // runtime stride
}
// The prologue's scratch (7 bytes per cell of the WHOLE pouch: row lengths, slots, copy masks, parts, diagonal
// indices) lies over the two buffers, which are initialised after it is dead; for a cluster CTA of a big pouch it
// can be larger than the buffers (n_total = cells of the pouch; 0: not more than the slots, single CTA).
static inline size_t integrate_scratch_bytes(int n_total) { return ((size_t)7 * ((n_total + 31) & ~31) + 127) & ~(size_t)127; }
static inline size_t integrate_smem_bytes(int np, int nslot, size_t elem2, int copies, bool unit, int hmax = 0, int n_total = 0)
{
    const size_t bufs = 2 * integrate_buf_stride(np, nslot, elem2, copies, unit, hmax);
    const size_t scr = integrate_scratch_bytes(n_total);
    return (bufs > scr ? bufs : scr) + integrate_tail_bytes(nslot, elem2, unit);
}

// One Euler step for one cell, reference source order (core/pouch.py:77-102), every operation
// individually rounded.  lap_c / lap_p are the already D-scaled Laplacian terms (:346-347).
template <typename R>
struct StepConsts {
    R k_1, k_a, k_p, k_2, V_S, KS2, KP2, K_5, c_tot, k_tau_4, k_i, D_c, D_p, dt;
    typename Strict<R>::Recip beta, rec_den, k_i_r;   // divisors that never change: reciprocal hoisted
};

// The part of a step that needs only the cell's own state (core/pouch.py:77-85, 91, 98-102): the
// three non-constant fluxes and the complete r update.  FP64 heavy, no shared memory.
template <typename R>
__device__ __forceinline__ void step_local(const StepConsts<R> &q, R ca, R ipt, R sv, R rv, R vplc,
                                           R &J_ch, R &J_S, R &J_PLC, R &r_new)
{
    typedef Strict<R> S;
    const R num = S::mul(S::mul(rv, ca), ipt);                              // :77
    const R den = S::mul(S::add(q.k_a, ca), S::add(q.k_p, ipt));            // :78-80
    const R x = S::fdiv(num, den);                                          // :80
    const R x3 = S::mul(S::mul(x, x), x);                                   // x**3 as numba expands it
    J_ch = S::mul(S::add(S::mul(q.k_1, x3), q.k_2), S::sub(sv, ca));        // :81
    const R ca_sq = S::mul(ca, ca);                                         // :84
    J_S = S::fdiv(S::mul(q.V_S, ca_sq), S::add(ca_sq, q.KS2));              // :85
    J_PLC = S::fdiv(S::mul(vplc, ca_sq), S::add(ca_sq, q.KP2));             // :91
    const R ca_4 = S::mul(ca_sq, ca_sq);                                    // :98
    const R rec = S::div(S::add(q.k_tau_4, ca_4), q.rec_den);               // :100
    const R inact = S::sub((R)1, S::div(S::mul(rv, S::add(q.k_i, ca)), q.k_i_r));         // :101
    r_new = S::add(rv, S::mul(S::mul(q.dt, rec), inact));                   // :102
}

// The part that needs the neighbours (core/pouch.py:87, 92): lap_c / lap_p are the D-scaled rows.
template <typename R>
__device__ __forceinline__ void step_finish(const StepConsts<R> &q, R ca, R ipt, R lap_c, R lap_p,
                                            R J_ch, R J_S, R J_PLC, R &c_new, R &p_new)
{
    typedef Strict<R> S;
    c_new = S::add(ca, S::mul(q.dt, S::sub(S::add(lap_c, J_ch), J_S)));     // :87
    p_new = S::add(ipt, S::mul(q.dt, S::sub(S::add(lap_p, J_PLC), S::mul(q.K_5, ipt))));  // :92
}

// ---- split-phase CTA barrier (mbarrier): arrive after publishing, wait before gathering -----------
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
                 "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(
                     (unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    } while (!ok);
}

// ---- thread-block-cluster helpers (distributed shared memory) --------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Local arrival that also announces `bytes` of asynchronous remote stores (st.async complete_tx) the
// barrier's current phase has to wait for.
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(
                     (unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
// 16-byte (two-real) shared-memory accesses at [32-bit shared address + compile-time immediate]: the
// immediate carries the step-parity buffer offset, so a gather is exactly one LDS per entry.
template <typename R> struct Smem2;
template <> struct Smem2<double> {
    template <int IMM> static __device__ __forceinline__ double2 ld(uint32_t addr)
    {
        double2 v;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(addr), "n"(IMM) : "memory");
        return v;
    }
    template <int IMM> static __device__ __forceinline__ void st(uint32_t addr, double x, double y)
    {
        asm volatile("st.shared.v2.f64 [%0+%3], {%1, %2};" ::"r"(addr), "d"(x), "d"(y), "n"(IMM) : "memory");
    }
    template <int IMM> static __device__ __forceinline__ void st_cluster(uint32_t addr, double x, double y)
    {
        asm volatile("st.shared::cluster.v2.f64 [%0+%3], {%1, %2};" ::"r"(addr), "d"(x), "d"(y), "n"(IMM) : "memory");
    }
    // asynchronous remote store that completes 16 transaction bytes on the destination CTA's mbarrier
    template <int IMM> static __device__ __forceinline__ void st_async(uint32_t addr, double x, double y, uint32_t mbar)
    {
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0+%4], {%1, %2}, [%3];"
                     ::"r"(addr), "d"(x), "d"(y), "r"(mbar), "n"(IMM) : "memory");
    }
};
template <> struct Smem2<float> {
    template <int IMM> static __device__ __forceinline__ float2 ld(uint32_t addr)
    {
        float2 v;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(v.x), "=f"(v.y) : "r"(addr), "n"(IMM) : "memory");
        return v;
    }
    template <int IMM> static __device__ __forceinline__ void st(uint32_t addr, float x, float y)
    {
        asm volatile("st.shared.v2.f32 [%0+%3], {%1, %2};" ::"r"(addr), "f"(x), "f"(y), "n"(IMM) : "memory");
    }
    template <int IMM> static __device__ __forceinline__ void st_cluster(uint32_t addr, float x, float y)
    {
        asm volatile("st.shared::cluster.v2.f32 [%0+%3], {%1, %2};" ::"r"(addr), "f"(x), "f"(y), "n"(IMM) : "memory");
    }
    template <int IMM> static __device__ __forceinline__ void st_async(uint32_t addr, float x, float y, uint32_t mbar)
    {
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0+%4], {%1, %2}, [%3];"
                     ::"r"(addr), "f"(x), "f"(y), "r"(mbar), "n"(IMM) : "memory");
    }
};

// Entries per region and bytes between the two step-parity buffers.  When both buffers fit with
// NSLOT-sized regions the stride is a compile-time constant (CSTRIDE): parity becomes an immediate
// and slots beyond n own real (unused) entries, so nothing in the step loop is predicated on them.
template <typename R, int NSLOT, int COPIES, bool UNIT, int HMAX = 0>
struct BufGeom {
    static constexpr size_t ESZ = 2 * sizeof(R);
    static constexpr int DIAGR = 0;                       // (round 1 kept a region of diagonal entries here)
    static constexpr size_t STRIDE = ((((size_t)(COPIES + DIAGR) * NSLOT + 1 + HMAX) * ESZ) + 127) & ~(size_t)127;
    static constexpr size_t TAIL = 2 * (size_t)NSLOT + 64;
    static constexpr bool CSTRIDE = 2 * STRIDE + TAIL <= 227 * 1024;
};

// Halo entries a cluster CTA can hold (values of other CTAs' cells that its rows read).
constexpr int CAB_HMAX = 384;

// EP = number of packed entry-index pairs kept in registers per cell (row length <= 2*EP).
// UNIT: every off-diagonal value is exactly 1.0 (verified in the prologue).  !UNIT: general weights,
// read from global/L1 each step and multiplied per entry.
// NCTA > 1: the pouch is split over a thread-block cluster of NCTA CTAs (cab200_plan_cluster): each
// CTA owns one part of the cells; entries of other parts are read from local *halo* entries which
// the owner pushes through distributed shared memory when it publishes -- with st.async, whose
// completion is counted in transaction bytes on the destination CTA's step barrier.  A CTA's
// barrier phase therefore completes when its own warps have published AND every halo value of that
// time has landed: no remote arrivals and no cluster-scope fences in the step loop.  Two barriers
// alternate with the step parity because a peer may already push time t+1 while this CTA's phase t
// is still pending (never further ahead: its push of t+2 needs this CTA's pushes of t+1, which are
// issued after phase t completed here).  Overwriting a halo entry of buffer t&1 with time t+2 is
// safe for the same reason: every warp that reads halo entries of a peer also owns cells pushed to
// that peer, and the peer's phase t+1 waits for those pushes, issued after the warp's gather of t.
template <typename R, int THREADS, int CPT, int EP, bool UNIT, int COPIES, int MINB, int NCTA = 1>
__global__ void __launch_bounds__(THREADS, MINB) integrate_kernel(const SimArgs a)
{
    typedef Strict<R> S;
    typedef typename S::vec2 R2;
    typedef Smem2<R> M;
    constexpr int NSLOT = THREADS * CPT;
    constexpr int EMAX = 2 * EP;
    constexpr int HMAX = (NCTA > 1) ? CAB_HMAX : 0;
    constexpr int DIAGR = 0;
    typedef BufGeom<R, NSLOT, COPIES, UNIT, HMAX> G;
    static_assert(NCTA == 1 || (G::CSTRIDE && UNIT), "cluster kernels need the compile-time layout");
    constexpr bool CSTRIDE = G::CSTRIDE;
    constexpr uint32_t ESZ = (uint32_t)G::ESZ;
    constexpr int STRIDE_IMM = CSTRIDE ? (int)G::STRIDE : 0;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int n = a.n, np = a.np;
    const int npr = CSTRIDE ? NSLOT : np;                 // entries per value region
    const uint32_t stride = CSTRIDE ? (uint32_t)G::STRIDE
                                    : (uint32_t)(((((size_t)(COPIES + DIAGR) * np + 1) * ESZ) + 127) & ~(size_t)127);
    // persistent table behind the buffers (and behind the prologue scratch, which may be the larger of the two in a
    // cluster CTA: integrate_scratch_bytes)
    const size_t scr_bytes = (NCTA > 1) ? (((size_t)7 * np + 127) & ~(size_t)127) : 0;
    const size_t tail_off = 2 * (size_t)stride > scr_bytes ? 2 * (size_t)stride : scr_bytes;
    uint16_t *orig_of = reinterpret_cast<uint16_t *>(smem_raw + tail_off);   // [NSLOT]
    // prologue scratch lies over the two buffers (initialised after the scratch is dead)
    unsigned char *scr = smem_raw;
    uint8_t *rlen = scr;                                                           // [n]
    uint16_t *slot_of = reinterpret_cast<uint16_t *>(scr + np);                    // [n]
    uint16_t *copy_sel = slot_of + np;                                             // [n]
    uint8_t *part_of = reinterpret_cast<uint8_t *>(copy_sel + np);                 // [n] (clusters)
    int8_t *rdg = reinterpret_cast<int8_t *>(part_of + np);                        // [n] diagonal index in the row, -1, -2

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int pouch = a.pouch_list[blockIdx.x / NCTA];
    const int64_t coff = a.cell_off[pouch];
    const uint32_t rank = (NCTA > 1) ? cluster_ctarank() : 0u;
    int bad = 0;

    // ---- prologue 1: row lengths, where the diagonal sits, which rows take the legacy layout ----
    for (int i = tid; i < n; i += THREADS) {
        const int rb = a.rowptr[i];
        int len = a.rowptr[i + 1] - rb;
        if (len > EMAX || len < 0) { bad = 1; len = 0; }
        rlen[i] = (uint8_t)len;
        int nd = 0, di = -1;
        for (int j = 0; j < len; ++j)
            if (a.colidx[rb + j] == i) { ++nd; di = j; }
        rdg[i] = (int8_t)(nd == 0 ? -1 : (nd == 1 ? di : -2));
    }
    for (int s = tid; s < NSLOT; s += THREADS) orig_of[s] = 0xffff;
    __syncthreads();
    // ---- prologue 2: cell -> slot.  Host plan if given, else stable rank by (length desc, id asc) --
    if (NCTA > 1) {
        // cluster plan: this CTA owns the cells of part `rank`; slot_of is the slot inside the part
        for (int i = tid; i < n; i += THREADS) {
            const int pq = a.slot_plan[i];
            int sl = a.slot_plan[n + i];
            if (pq >= NCTA || sl >= NSLOT) { bad = 1; sl = 0; }
            part_of[i] = (uint8_t)pq;
            slot_of[i] = (uint16_t)sl;
            copy_sel[i] = (COPIES == 2) ? a.slot_plan[2 * n + i] : (uint16_t)0;
            if (pq == (int)rank) orig_of[sl] = (uint16_t)i;
        }
        __syncthreads();
        for (int i = tid; i < n; i += THREADS)
            if (part_of[i] == rank && orig_of[slot_of[i]] != i) bad = 1;
    } else if (a.slot_plan) {
        for (int i = tid; i < n; i += THREADS) {
            int sl = a.slot_plan[i];
            if (sl >= n) { bad = 1; sl = i; }
            slot_of[i] = (uint16_t)sl;
            orig_of[sl] = (uint16_t)i;
            copy_sel[i] = (COPIES == 2) ? a.slot_plan[n + i] : (uint16_t)0;
        }
        __syncthreads();
        for (int i = tid; i < n; i += THREADS)       // a permutation hits every slot exactly once
            if (orig_of[slot_of[i]] != i) bad = 1;
    } else {
        for (int i = tid; i < n; i += THREADS) {
            const int li = rlen[i];
            int rank = 0;
            for (int j = 0; j < n; ++j) {
                const int lj = rlen[j];
                rank += (lj > li) || (lj == li && j < i);
            }
            slot_of[i] = (uint16_t)rank;
            orig_of[rank] = (uint16_t)i;
            copy_sel[i] = 0;
        }
    }
    __syncthreads();

    // ---- prologue 3: parameters ------------------------------------------------------------------
    const double *prm = a.params + (size_t)pouch * CAB200_NPARAM;
    StepConsts<R> q;
    {
        const R K_S = (R)prm[CAB200_KSERCA], K_PLC = (R)prm[CAB200_KPLC];
        const R k_tau = (R)prm[CAB200_KTAU], tau_max = (R)prm[CAB200_TAUMAX];
        q.k_1 = (R)prm[CAB200_K1]; q.k_a = (R)prm[CAB200_KA]; q.k_p = (R)prm[CAB200_KP];
        q.k_2 = (R)prm[CAB200_K2]; q.V_S = (R)prm[CAB200_VSERCA]; q.K_5 = (R)prm[CAB200_K5];
        q.c_tot = (R)prm[CAB200_CTOT]; q.k_i = (R)prm[CAB200_KI];
        q.beta = S::recip((R)prm[CAB200_BETA]); q.k_i_r = S::recip(q.k_i);
        q.D_c = (R)prm[CAB200_DC]; q.D_p = (R)prm[CAB200_DP]; q.dt = (R)prm[CAB200_DT];
        q.KS2 = S::mul(K_S, K_S);                                              // :85
        q.KP2 = S::mul(K_PLC, K_PLC);                                          // :91
        q.k_tau_4 = S::mul(S::mul(S::mul(k_tau, k_tau), k_tau), k_tau);        // :99
        q.rec_den = S::recip(S::mul(tau_max, q.k_tau_4));                      // :100
    }

    // ---- prologue 4: private per-cell state ------------------------------------------------------
    // Slot of (warp w, group k, lane l) = (w*CPT + k)*32 + l: the host planner deals cells with similar
    // row shapes to the same warp, so ONE warp-uniform position range [start, end) serves all CPT*32
    // cells of a warp.  k1_row (k1_layout.cuh) places a row's entries among the EMAX positions: entries
    // before the diagonal right-aligned below the warp's diagonal point D, the others left-aligned from D;
    // the diagonal itself is added from registers inside the pair `dpair`.  Everything else points at the zero entry; the
    // step loop enters the fully unrolled position sequence at `start` through one jump and leaves it
    // at `end`.  ent[k][pos] is the 32-bit shared address of the entry inside buffer 0; buffer 1 is
    // +stride.
    const uint32_t sb0 = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t zero_addr = sb0 + (uint32_t)((COPIES + DIAGR) * npr) * ESZ;
    const uint32_t halo_addr = zero_addr + ESZ;
    R c[CPT], p[CPT], s[CPT], r[CPT], vplc[CPT], dcoef[CPT];
    uint32_t ent[CPT][EMAX];
    uint32_t own[CPT];              // shared address (buffer 0) of the cell's own value entry
    int rowbeg[CPT], len[CPT];
    bool live[CPT];
    K1Row row[CPT];
    int min_first = EMAX, max_last = 0;
    // the warp's diagonal point (k1_layout.cuh): one D for the CPT*32 cells of the warp, from the largest number of
    // entries before / after the diagonal among them -- the planner has made sure these go together
    int max_nb = 0, max_na = 0;
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int cell = (int)orig_of[(warp * CPT + k) * 32 + lane];
        if (cell != 0xffff) {
            max_nb = max(max_nb, k1_row_nb(rlen[cell], rdg[cell]));
            max_na = max(max_na, k1_row_na(rlen[cell], rdg[cell]));
        }
    }
    int dpt = UNIT ? k1_warp_diag(EP, __reduce_max_sync(0xffffffffu, max_nb), __reduce_max_sync(0xffffffffu, max_na)) : EP;
    if (dpt < 0) { bad = 1; dpt = EP; }
    const int dpair = k1_diag_pair(EP, dpt);          // warp-uniform: the pair of positions that carries the diagonal
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int slot = (warp * CPT + k) * 32 + lane;
        const int cell = (int)orig_of[slot];          // 0xffff: no cell in this slot
        live[k] = (cell != 0xffff);
        own[k] = sb0 + (uint32_t)slot * ESZ;
        rowbeg[k] = 0; len[k] = 0;
        c[k] = p[k] = s[k] = r[k] = vplc[k] = dcoef[k] = (R)0;
        row[k] = k1_row(UNIT, EP, 0, -1, dpt);
        if (live[k]) {
            len[k] = rlen[cell];
            rowbeg[k] = a.rowptr[cell];
            row[k] = k1_row(UNIT, EP, len[k], rdg[cell], dpt);
            c[k] = (R)a.c0[coff + cell];
            p[k] = (R)a.p0[coff + cell];
            r[k] = (R)a.r0[coff + cell];
            vplc[k] = (R)a.vplc[coff + cell];
            if (len[k] > 0) { min_first = min(min_first, row[k].first); max_last = max(max_last, row[k].last); }
            if (UNIT && row[k].legacy) bad = 1;       // several diagonal entries: this geometry needs the general kernel
        }
        s[k] = S::div(S::sub(q.c_tot, c[k]), q.beta);                          // core/pouch.py:322
    }
    // cluster: halo lists (sorted global cell ids per part) follow the three per-cell arrays
    const uint16_t *hcount = a.slot_plan + 3 * (size_t)n;
    const uint16_t *hlists = hcount + NCTA;
    auto halo_find = [&](int part, int cell_id) -> int {      // index of cell_id in part's halo list, or -1
        int base = 0;
        for (int qq = 0; qq < part; ++qq) base += hcount[qq];
        int lo = 0, hi = (int)hcount[part] - 1;
        while (lo <= hi) {
            const int mid = (lo + hi) >> 1;
            const int v = hlists[base + mid];
            if (v == cell_id) return mid;
            if (v < cell_id) lo = mid + 1; else hi = mid - 1;
        }
        return -1;
    };
    int start, end;
    k1_warp_range(UNIT, EP, __reduce_min_sync(0xffffffffu, min_first), __reduce_max_sync(0xffffffffu, max_last),
                  dpt, &start, &end);
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int slot = (warp * CPT + k) * 32 + lane;
        const int cell = live[k] ? (int)orig_of[slot] : 0;
        const uint32_t sel = live[k] ? (uint32_t)copy_sel[cell] : 0u;
        if (UNIT && live[k] && row[k].skip)                                   // register diagonal: its weight
            dcoef[k] = (R)a.vals[rowbeg[k] + row[k].split];
#pragma unroll
        for (int pos = 0; pos < EMAX; ++pos) {
            const int e = live[k] ? k1_entry_at(row[k], EP, pos) : -1;
            uint32_t o = zero_addr;
            if (e >= 0 && e < len[k]) {
                const int col = a.colidx[rowbeg[k] + e];
                if (col < 0 || col >= n) bad = 1;
                else if (NCTA > 1 && part_of[col] != rank) {
                    const int h = halo_find((int)rank, col);        // another CTA's cell: local halo entry
                    if (h < 0 || h >= HMAX || a.vals[rowbeg[k] + e] != 1.0) bad = 1;
                    else o = halo_addr + (uint32_t)h * ESZ;
                } else {
                    uint32_t j = slot_of[col];
                    if (COPIES == 2 && ((sel >> e) & 1u)) j = (uint32_t)npr + (j ^ 4u);
                    o = sb0 + j * ESZ;
                    if (UNIT) {
                        const double w = a.vals[rowbeg[k] + e];
                        if (col == cell || w != 1.0) bad = 1;       // the diagonal is never gathered (row[k].skip)
                    }
                }
            }
            ent[k][pos] = o;
        }
    }
    // cluster: where this thread's cells must be pushed (their halo entries in the other CTAs)
    uint32_t push[CPT][NCTA > 1 ? NCTA - 1 : 1];
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
#pragma unroll
        for (int d = 0; d < (NCTA > 1 ? NCTA - 1 : 1); ++d) push[k][d] = 0u;
        if (NCTA > 1 && live[k]) {
            const int cell = (int)orig_of[(warp * CPT + k) * 32 + lane];
#pragma unroll
            for (int d = 0; d < NCTA - 1; ++d) {
                const uint32_t q2 = (rank + 1 + d) % NCTA;
                const int h = halo_find((int)q2, cell);
                if (h >= HMAX) bad = 1;
                else if (h >= 0)
                    push[k][d] = mapa_shared(halo_addr + (uint32_t)h * ESZ, q2);
            }
        }
    }
    if (bad) atomicOr(a.status_out + pouch, CAB200_ST_BADGEOM);
#ifdef CAB_K1_CHECK
    unsigned chk_adj_local = 0u;           // ranks whose cells sit in this CTA's halo list, i.e. who push to this CTA
    if (NCTA > 1) {
        int hbase = 0;
        for (int qq = 0; qq < (int)rank; ++qq) hbase += hcount[qq];
        for (int h = tid; h < (int)hcount[rank]; h += THREADS) chk_adj_local |= 1u << part_of[hlists[hbase + h]];
    }
#endif
    __syncthreads();   // scratch (inside buffer 1) is dead from here on
    for (uint32_t i = tid; i < 2 * stride / 16; i += THREADS)
        reinterpret_cast<uint4 *>(smem_raw)[i] = make_uint4(0u, 0u, 0u, 0u);
    __shared__ uint64_t step_bar[2];     // clusters alternate between two (see above); else [0] only
#ifdef CAB_K1_CHECK
    // Debug build (compute-sanitizer's racecheck is not available on every pool): the buffer-reuse argument of the
    // step loop as run-time checks.  ann_time[w] = the time warp w is ABOUT to publish (written before its stores and
    // pushes), pub_time[w] = the time it HAS published (written after them, before its arrival).  A warp that gathers
    // time t must find every local pub_time >= t (the values are there) and every ann_time <= t + 1 -- locally and in
    // every peer CTA that pushes halo values to it (nobody has started to overwrite buffer t&1 with time t + 2).
    // A violation sets CAB200_ST_PROTOCOL in the pouch's status word.
    __shared__ int ann_time[32], pub_time[32], chk_adj;
    if (tid < 32) { ann_time[tid] = 0; pub_time[tid] = 0; }
    if (tid == 0) chk_adj = 0;
#endif
    if (tid == 0) {
        mbar_init(&step_bar[0], THREADS / 32);         // one arrival per warp
        mbar_init(&step_bar[1], THREADS / 32);
        if (NCTA > 1) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
#ifdef CAB_K1_CHECK
    if (chk_adj_local) atomicOr(&chk_adj, (int)chk_adj_local);
    __syncthreads();
    const unsigned chk_peers = (unsigned)chk_adj & ~(1u << rank);
    const uint32_t ann_addr = (uint32_t)__cvta_generic_to_shared(&ann_time[0]);
#define CAB_CHECK_ANNOUNCE(TNEW) if (lane == 0) ((volatile int *)ann_time)[warp] = (TNEW);
#define CAB_CHECK_PUBLISHED(TNEW) if (lane == 0) ((volatile int *)pub_time)[warp] = (TNEW);
#define CAB_CHECK_GATHERED(TCUR)                                                                   \
    {                                                                                             \
        int viol_ = 0;                                                                            \
        if (lane < THREADS / 32) {                                                                \
            const int pv_ = ((volatile int *)pub_time)[lane], av_ = ((volatile int *)ann_time)[lane]; \
            viol_ |= (pv_ < (TCUR)) || (av_ > (TCUR) + 1);                                        \
            if (NCTA > 1) {                                                                       \
                _Pragma("unroll") for (int q_ = 0; q_ < NCTA; ++q_)                               \
                    if ((chk_peers >> q_) & 1u) {                                                 \
                        int rv_;                                                                  \
                        asm volatile("ld.volatile.shared::cluster.u32 %0, [%1];" : "=r"(rv_)      \
                                     : "r"(mapa_shared(ann_addr + 4u * (uint32_t)lane, (uint32_t)q_)) : "memory"); \
                        viol_ |= rv_ > (TCUR) + 1;                                                \
                    }                                                                             \
            }                                                                                     \
        }                                                                                         \
        if (viol_) atomicOr(a.status_out + pouch, CAB200_ST_PROTOCOL);                            \
    }
#else
#define CAB_CHECK_ANNOUNCE(TNEW)
#define CAB_CHECK_PUBLISHED(TNEW)
#define CAB_CHECK_GATHERED(TCUR)
#endif
    if (NCTA > 1) cluster_sync_all();    // every CTA's buffers are zeroed and its barrier exists

    // cluster: the peers' step barriers (as cluster addresses) that this thread's pushes complete on
    const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(&step_bar[0]);
    uint32_t bar_peer[NCTA > 1 ? NCTA - 1 : 1];
#pragma unroll
    for (int d = 0; d < (NCTA > 1 ? NCTA - 1 : 1); ++d)
        bar_peer[d] = (NCTA > 1) ? mapa_shared(bar_addr, (rank + 1 + d) % NCTA) : bar_addr;
    const uint32_t copyb_off = (uint32_t)npr * ESZ;                 // second value copy (multiple of 128 B)
    // Publishes a cell's new (c, p): value entry and second copy in the bank group 4 away.  IMM selects
    // the step-parity buffer at compile time (CSTRIDE) -- else roff does.
#define CAB_PUBLISH(IMM, ROFF, K, CV, PV, ASYNC_BAR)                                               \
    if (CSTRIDE || live[K]) {                                                                     \
        M::template st<IMM>(own[K] + (ROFF), (CV), (PV));                                         \
        if (COPIES == 2) M::template st<IMM>(((own[K] + copyb_off) ^ (4u * ESZ)) + (ROFF), (CV), (PV)); \
        if (NCTA > 1) {                                                                           \
            _Pragma("unroll") for (int d_ = 0; d_ < NCTA - 1; ++d_)                               \
                if (push[K][d_]) {                                                                \
                    if ((ASYNC_BAR) < 0) M::template st_cluster<IMM>(push[K][d_], (CV), (PV));    \
                    else M::template st_async<IMM>(push[K][d_], (CV), (PV), bar_peer[d_] + 8u * (ASYNC_BAR)); \
                }                                                                                 \
        }                                                                                         \
    }
#pragma unroll
    for (int k = 0; k < CPT; ++k) { CAB_PUBLISH(0, 0u, k, c[k], p[k], -1) }
    __syncthreads();
    if (NCTA > 1) cluster_sync_all();    // initial values, halo entries included, are in place
    const uint32_t halo_bytes = (NCTA > 1) ? (uint32_t)hcount[rank] * ESZ : 0u;   // pushed to this CTA per step
    // lane 0 of a warp: arrival on the barrier that guards the values of the NEXT time (index BAR);
    // warp 0 also announces the halo bytes that phase has to receive
    auto arrive_step = [&](int bar) {
        if (NCTA > 1 && warp == 0) mbar_arrive_expect_tx(&step_bar[bar], halo_bytes);
        else mbar_arrive(&step_bar[bar]);
    };

    R *frames = reinterpret_cast<R *>(a.frames_out) + (size_t)coff * 4 * a.n_frames;
    int fi = 0;
    int cap_t = (a.n_frames > 0) ? a.frame_steps[0] : -1;
    const int T = a.T;

    // Writes the state at the current time (c, p, s, r of every owned cell) as frame fi: in original
    // cell order (scattered 8-byte stores; fine for a few frames), or -- frames_tmp -- in slot order,
    // one fully coalesced store per state and slot group.
    R *ftmp = a.frames_tmp ? reinterpret_cast<R *>(a.frames_tmp) + (size_t)(blockIdx.x / NCTA) * a.n_frames * 4 * (NCTA * NSLOT)
                                 + (size_t)rank * NSLOT + (size_t)warp * CPT * 32 + lane
                           : nullptr;
    auto capture = [&]() {
        int nonfinite = 0;
        if (ftmp) {
            R *f = ftmp + (size_t)fi * 4 * (NCTA * NSLOT);
#pragma unroll
            for (int k = 0; k < CPT; ++k) {
                f[k * 32] = c[k]; f[k * 32 + NCTA * NSLOT] = p[k];
                f[k * 32 + 2 * (NCTA * NSLOT)] = s[k]; f[k * 32 + 3 * (NCTA * NSLOT)] = r[k];
                nonfinite |= !(isfinite(c[k]) && isfinite(p[k]) && isfinite(s[k]) && isfinite(r[k]));
            }
        } else {
#pragma unroll
            for (int k = 0; k < CPT; ++k) {
                if (live[k]) {
                    const int cell = orig_of[(warp * CPT + k) * 32 + lane];
                    R *f = frames + (size_t)fi * 4 * n + cell;
                    f[0] = c[k]; f[n] = p[k]; f[2 * (size_t)n] = s[k]; f[3 * (size_t)n] = r[k];
                    nonfinite |= !(isfinite(c[k]) && isfinite(p[k]) && isfinite(s[k]) && isfinite(r[k]));
                }
            }
        }
        if (nonfinite) atomicOr(a.status_out + pouch, CAB200_ST_NONFINITE);
        ++fi;
        cap_t = (fi < a.n_frames) ? a.frame_steps[fi] : -1;
    };

    // Step organisation.  Everything that needs only the cell's own state (step_local: the five
    // divisions, ~75 % of the FP64 work) is computed one step ahead, *between* publishing the new
    // (c, p) and waiting for everybody else's: publish -> arrive -> local work -> wait -> gather ->
    // finish -> publish ...  The barrier is split (mbarrier arrive / try_wait), so a warp rarely
    // blocks.  Buffer safety: a thread overwrites buffer (t+1)&1 only after passing wait(t), i.e.
    // after every warp has arrived for step t, which each does after its own gather of step t-1 --
    // the last reader of that buffer.  Slots beyond n carry harmless zeros (no NaNs arise).
    if (cap_t == 0) capture();
    R J_ch[CPT], J_S[CPT], J_PLC[CPT];
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        R rn;
        step_local<R>(q, c[k], p[k], s[k], r[k], vplc[k], J_ch[k], J_S[k], J_PLC[k], rn);
        r[k] = rn;      // r now holds r(t+1); r(t) was captured above if requested
    }
    if (lane == 0) mbar_arrive(&step_bar[0]);   // state(0) was published (and pushed) before the barriers above
    unsigned parity = 0;

#ifdef CAB_X_NO_EXIT
#define CAB_EXIT_CHECK(P0)
#else
#define CAB_EXIT_CHECK(P0) if (end <= (P0) + 2) break;
#endif
// The register diagonal of every owned cell: L_ii * (c_i, p_i), each product rounded like a stored diagonal
// entry would be (dcoef is 0 for empty slots and rows without a diagonal: +-0 never changes a sum that cannot
// be -0).
#define CAB_DIAG_ADD                                                                              \
    _Pragma("unroll") for (int k = 0; k < CPT; ++k) {                                             \
        sc[k] = S::add(sc[k], S::mul(dcoef[k], c[k])); sp[k] = S::add(sp[k], S::mul(dcoef[k], p[k])); \
    }

// One pair of row positions for every owned cell: all loads of the pair first (2*CPT independent
// 16-byte loads in flight per thread), then the additions in position order.  The warp's pair `dpair` also
// carries the register diagonal at its place in the order (a warp-uniform test per pair).
#define CAB_GATHER_PAIR(IMM, ROFF, P0)                                                            \
    {                                                                                             \
        R2 va[CPT], vb[CPT];                                                                      \
        _Pragma("unroll") for (int k = 0; k < CPT; ++k) {                                         \
            va[k] = M::template ld<IMM>(ent[k][(P0)] + (ROFF));                                   \
            vb[k] = M::template ld<IMM>(ent[k][(P0) + 1] + (ROFF));                               \
        }                                                                                         \
        if (UNIT) {                                                                               \
            _Pragma("unroll") for (int k = 0; k < CPT; ++k) {                                     \
                sc[k] = S::add(sc[k], va[k].x); sp[k] = S::add(sp[k], va[k].y);                   \
            }                                                                                     \
            if ((EP & 1) && dpair == (P0)) { CAB_DIAG_ADD }                                       \
            _Pragma("unroll") for (int k = 0; k < CPT; ++k) {                                     \
                sc[k] = S::add(sc[k], vb[k].x); sp[k] = S::add(sp[k], vb[k].y);                   \
            }                                                                                     \
            if (!(EP & 1) && dpair == (P0)) { CAB_DIAG_ADD }                                      \
        } else {                                                                                  \
            _Pragma("unroll") for (int k = 0; k < CPT; ++k) {                                     \
                const int e0 = (P0) - row[k].lo_base;                                             \
                const R w0 = (e0 >= 0 && e0 < len[k]) ? (R)__ldg(a.vals + rowbeg[k] + e0) : (R)0;  \
                const R w1 = (e0 + 1 >= 0 && e0 + 1 < len[k]) ? (R)__ldg(a.vals + rowbeg[k] + e0 + 1) : (R)0; \
                sc[k] = S::add(sc[k], S::mul(w0, va[k].x)); sp[k] = S::add(sp[k], S::mul(w0, va[k].y)); \
                sc[k] = S::add(sc[k], S::mul(w1, vb[k].x)); sp[k] = S::add(sp[k], S::mul(w1, vb[k].y)); \
            }                                                                                     \
        }                                                                                         \
        CAB_EXIT_CHECK(P0)                                                                        \
    }

// One time step reading the buffer at +CUR and writing the one at +NXT (immediates when CSTRIDE).
#define CAB_STEP(CUR_IMM, NXT_IMM, CUR_R, NXT_R, CUR_BAR, NXT_BAR)                                \
    {                                                                                             \
        mbar_wait(&step_bar[CUR_BAR], parity);                                                    \
        if (NCTA == 1 || (CUR_BAR) == 1) parity ^= 1u;   /* a cluster uses each barrier every 2nd step */ \
        R sc[CPT], sp[CPT];                                                                       \
        _Pragma("unroll") for (int k = 0; k < CPT; ++k) { sc[k] = (R)0; sp[k] = (R)0; }           \
        switch (start) { /* warp-uniform entry point; later positions follow until `end` */       \
            case 0: CAB_GATHER_PAIR(CUR_IMM, CUR_R, 0)                                            \
            case 2: CAB_GATHER_PAIR(CUR_IMM, CUR_R, 2)                                            \
            case 4: CAB_GATHER_PAIR(CUR_IMM, CUR_R, 4)                                            \
            case 6: CAB_GATHER_PAIR(CUR_IMM, CUR_R, 6)                                            \
            case 8: CAB_GATHER_PAIR(CUR_IMM, CUR_R, 8)                                            \
            case 10: if constexpr (EMAX > 10) CAB_GATHER_PAIR(CUR_IMM, CUR_R, 10 % EMAX)          \
            case 12: if constexpr (EMAX > 12) CAB_GATHER_PAIR(CUR_IMM, CUR_R, 12 % EMAX)          \
            case 14: if constexpr (EMAX > 14) CAB_GATHER_PAIR(CUR_IMM, CUR_R, 14 % EMAX)          \
            default: break;                                                                       \
        }                                                                                         \
        CAB_CHECK_GATHERED(t)                                                                     \
        CAB_CHECK_ANNOUNCE(t + 1)                                                                 \
        _Pragma("unroll") for (int k = 0; k < CPT; ++k) {                                         \
            const R lap_c = S::mul(q.D_c, sc[k]);                              /* :346 */         \
            const R lap_p = S::mul(q.D_p, sp[k]);                              /* :347 */         \
            R cn, pn;                                                                             \
            step_finish<R>(q, c[k], p[k], lap_c, lap_p, J_ch[k], J_S[k], J_PLC[k], cn, pn);       \
            c[k] = cn; p[k] = pn;                                                                 \
            CAB_PUBLISH(NXT_IMM, NXT_R, k, cn, pn, NXT_BAR)                                       \
        }                                                                                         \
        /* one arrival per warp: __syncwarp orders the other lanes' stores before lane 0's release */ \
        __syncwarp();                                                                             \
        CAB_CHECK_PUBLISHED(t + 1)                                                                \
        if (lane == 0) arrive_step(NXT_BAR);                                                      \
        /* own-state work for the next step, overlapping the other warps' gathers */              \
        _Pragma("unroll") for (int k = 0; k < CPT; ++k)                                           \
            s[k] = S::div(S::sub(q.c_tot, c[k]), q.beta);                      /* :95 */          \
        if (t + 1 == cap_t) capture();   /* uniform across the CTA; r[] holds r(t+1) here */       \
        _Pragma("unroll") for (int k = 0; k < CPT; ++k) {                                         \
            R rn;                                                                                 \
            step_local<R>(q, c[k], p[k], s[k], r[k], vplc[k], J_ch[k], J_S[k], J_PLC[k], rn);     \
            r[k] = rn;                                                                            \
        }                                                                                         \
    }

    const uint32_t rstride = CSTRIDE ? 0u : stride;     // runtime parity offset only without CSTRIDE
    int t = 0;
    constexpr int B1 = (NCTA > 1) ? 1 : 0;     // barrier guarding odd times (clusters only)
    for (; t + 1 < T - 1; t += 2) {
        CAB_STEP(0, STRIDE_IMM, 0u, rstride, 0, B1)
        ++t;
        CAB_STEP(STRIDE_IMM, 0, rstride, 0u, B1, 0)
        --t;
    }
    if (t < T - 1) { CAB_STEP(0, STRIDE_IMM, 0u, rstride, 0, B1) }
    if (NCTA > 1) {
        // the last pushes (time T-1) must have landed before anybody's shared memory goes away
        if (T > 1) mbar_wait(&step_bar[(T - 1) & 1], (unsigned)(((T - 1) >> 1) & 1));
        cluster_sync_all();
    }
#undef CAB_STEP
#undef CAB_CHECK_ANNOUNCE
#undef CAB_CHECK_PUBLISHED
#undef CAB_CHECK_GATHERED
#undef CAB_GATHER_PAIR
#undef CAB_DIAG_ADD
#undef CAB_EXIT_CHECK
#undef CAB_PUBLISH
}

}  // namespace cab200
