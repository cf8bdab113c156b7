// integrate.cuh -- K1, the persistent ensemble integrator (one CTA per pouch).
//
// Replaces reference core/pouch.py:338-366 (time loop), :31-104 (_simulate_time_step) and the two
// scipy CSR products per step (:346-347) for a whole ensemble in one launch.
//
// Data placement for one pouch (n cells), fixed for all T steps:
//   shared   two buffers (step parity) of ENT = (COPIES+1)*NP+1 entries, one entry = the pair (c, p):
//              [0, NP)            value entries   (c_i, p_i) at index slot_i
//              [NP, 2NP)          (COPIES == 2) a second copy of the values at NP + (slot_i ^ 4),
//                                 i.e. in the bank group 4 away: every gather may read either copy,
//                                 which the host planner uses to dodge bank conflicts
//              [COPIES*NP, +NP)   diagonal entries (L_ii*c_i, L_ii*p_i)
//              [(COPIES+1)*NP]    the zero entry (+0, +0), never written after the prologue
//            Step t reads buffer t&1 and writes buffer (t+1)&1: ONE __syncthreads() per step.
//   regs     per owned cell: c, p, s, r, V_PLC, L_ii and the row's entry indices (16-bit, packed in
//            pairs; time-invariant), plus the pouch's parameters (uniform registers).
//   HBM      only the captured frames, [F][4][n] in original cell order.
//
// A Laplacian row is then a branch-free gather: sum = +0; for each stored entry in ascending column
// order sum += entry -- off-diagonal entries index the neighbour's value entry (1.0*x == x exactly
// for the adjacency-minus-degree Laplacian), the diagonal entry indexes the cell's own diagonal
// entry, and rows shorter than the warp's longest row are padded with the zero entry.  Because the
// sum starts from +0.0 it can never be -0.0, so adding the +0.0 padding is an exact no-op and the
// result is bit-identical to scipy's csr_matvec followed by the reference's `D * (...)`.
//
// Cell -> thread ownership is static.  Cells are ranked by CSR row length (descending, stable) and
// dealt to slots `k*THREADS + tid`, so the 32 cells a warp handles together have (almost) equal
// row length and the padding costs next to nothing.
#pragma once
#include "common.cuh"

namespace cab200 {

struct SimArgs {
    const int32_t *pouch_list;  // [grid] pouch ids handled by this launch
    const int64_t *cell_off;    // [P+1]
    const int32_t *rowptr;      // this geometry's CSR
    const int32_t *colidx;
    const double *vals;
    const uint16_t *slot_plan;  // optional plan (cab200_plan_slots): slot_of[n] then copy_sel[n]
    int n;
    int np;                     // n rounded up to 32: entries per region
    const double *params;
    const double *c0, *p0, *r0, *vplc;
    int T, n_frames;
    const int32_t *frame_steps;
    void *frames_out;
    int32_t *status_out;
};

// Shared-memory bytes K1 needs for a pouch padded to `np` slots.
static inline size_t integrate_smem_bytes(int np, int nslot, size_t elem2, int copies)
{
    const size_t ent = ((size_t)copies + 1) * (size_t)np + 1;
    size_t bytes = 2 * ent * elem2;          // the two step-parity buffers
    bytes = (bytes + 15) & ~(size_t)15;
    bytes += 2 * (size_t)nslot;              // orig_of[NSLOT] u16
    return bytes;
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

// EP = number of packed entry-index pairs kept in registers per cell (row length <= 2*EP).
// UNIT: every off-diagonal value is exactly 1.0 (verified in the prologue).  !UNIT: general weights,
// read from global/L1 each step and multiplied per entry.
template <typename R, int THREADS, int CPT, int EP, bool UNIT, int COPIES, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) integrate_kernel(const SimArgs a)
{
    typedef Strict<R> S;
    typedef typename S::vec2 R2;
    constexpr int NSLOT = THREADS * CPT;
    constexpr int EMAX = 2 * EP;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = a.n, np = a.np;
    const int ent = (COPIES + 1) * np + 1;
    R2 *buf0 = reinterpret_cast<R2 *>(smem_raw);
    R2 *buf1 = buf0 + ent;
    uint16_t *orig_of = reinterpret_cast<uint16_t *>(
        smem_raw + ((2 * (size_t)ent * sizeof(R2) + 15) & ~(size_t)15));          // [NSLOT]
    // prologue scratch aliases buffer 1 (initialised after the scratch is dead)
    unsigned char *scr = reinterpret_cast<unsigned char *>(buf1);
    uint8_t *rlen = scr;                                                           // [n]
    uint16_t *slot_of = reinterpret_cast<uint16_t *>(scr + np);                    // [n]
    uint16_t *copy_sel = slot_of + np;                                             // [n]

    const int tid = threadIdx.x;
    const int pouch = a.pouch_list[blockIdx.x];
    const int64_t coff = a.cell_off[pouch];
    int bad = 0;

    // ---- prologue 1: row lengths -------------------------------------------------------------
    for (int i = tid; i < n; i += THREADS) {
        int len = a.rowptr[i + 1] - a.rowptr[i];
        if (len > EMAX || len < 0) { bad = 1; len = 0; }
        rlen[i] = (uint8_t)len;
    }
    for (int s = tid; s < NSLOT; s += THREADS) orig_of[s] = 0xffff;
    __syncthreads();
    // ---- prologue 2: cell -> slot.  Host plan if given, else stable rank by (length desc, id asc) --
    if (a.slot_plan) {
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

    // ---- prologue 3: parameters and private per-cell state -------------------------------------
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

    const uint32_t ZERO_ENT = (uint32_t)((COPIES + 1) * np);
    R c[CPT], p[CPT], s[CPT], r[CPT], vplc[CPT], dcoef[CPT];
    uint32_t nb[CPT][EP];
    int wlen[CPT], rowbeg[CPT];
    bool live[CPT];
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int slot = k * THREADS + tid;
        const int cell = (slot < n) ? (int)orig_of[slot] : 0xffff;
        live[k] = (cell != 0xffff);
        rowbeg[k] = 0;
        int len = 0;
        c[k] = p[k] = s[k] = r[k] = vplc[k] = dcoef[k] = (R)0;
#pragma unroll
        for (int e = 0; e < EP; ++e) nb[k][e] = ZERO_ENT | (ZERO_ENT << 16);
        if (live[k]) {
            len = rlen[cell];
            const uint32_t sel = copy_sel[cell];
            const int rb = a.rowptr[cell];
            rowbeg[k] = rb;
#pragma unroll
            for (int e = 0; e < EMAX; ++e) {
                if (e < len) {
                    const int col = a.colidx[rb + e];
                    uint32_t j = ZERO_ENT;
                    if (col < 0 || col >= n) bad = 1;
                    else {
                        j = slot_of[col];
                        if (COPIES == 2 && ((sel >> e) & 1u)) j = (uint32_t)np + (j ^ 4u);
                    }
                    if (UNIT) {
                        const double w = a.vals[rb + e];
                        if (col == cell) { dcoef[k] = (R)w; j = (uint32_t)(COPIES * np + slot); }
                        else if (w != 1.0) bad = 1;
                    }
                    const uint32_t sh = (e & 1) * 16;
                    nb[k][e >> 1] = (nb[k][e >> 1] & ~(0xffffu << sh)) | (j << sh);
                }
            }
            c[k] = (R)a.c0[coff + cell];
            p[k] = (R)a.p0[coff + cell];
            r[k] = (R)a.r0[coff + cell];
            vplc[k] = (R)a.vplc[coff + cell];
            s[k] = S::div(S::sub(q.c_tot, c[k]), q.beta);                      // core/pouch.py:322
        }
        wlen[k] = __reduce_max_sync(0xffffffffu, len);
    }
    if (bad) atomicOr(a.status_out + pouch, CAB200_ST_BADGEOM);
    __syncthreads();   // scratch (inside buffer 1) is dead from here on
    for (int i = tid; i < ent; i += THREADS) buf1[i] = S::make2((R)0, (R)0);
    for (int i = tid; i < ent; i += THREADS) buf0[i] = S::make2((R)0, (R)0);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        if (live[k]) {
            const int slot = k * THREADS + tid;
            buf0[slot] = S::make2(c[k], p[k]);
            if (COPIES == 2) buf0[np + (slot ^ 4)] = S::make2(c[k], p[k]);
            if (UNIT) buf0[COPIES * np + slot] = S::make2(S::mul(dcoef[k], c[k]), S::mul(dcoef[k], p[k]));
        }
    }
    __syncthreads();

    R *frames = reinterpret_cast<R *>(a.frames_out) + (size_t)coff * 4 * a.n_frames;
    int fi = 0;
    int cap_t = (a.n_frames > 0) ? a.frame_steps[0] : -1;
    const int T = a.T;
    const R2 *cur = buf0;
    R2 *nxt = buf1;

    // Writes the state at the current time (c, p, s, r of every owned cell) as frame fi.
    auto capture = [&]() {
        int nonfinite = 0;
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            if (live[k]) {
                const int cell = orig_of[k * THREADS + tid];
                R *f = frames + (size_t)fi * 4 * n + cell;
                f[0] = c[k]; f[n] = p[k]; f[2 * (size_t)n] = s[k]; f[3 * (size_t)n] = r[k];
                nonfinite |= !(isfinite(c[k]) && isfinite(p[k]) && isfinite(s[k]) && isfinite(r[k]));
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
    // blocks, warps drift apart, and the shared-memory-bound gathers of some warps overlap the
    // FP64-bound local work of others instead of all warps hammering one pipe at a time.
    // Buffer safety: a thread overwrites buffer (t+1)&1 only after passing wait(t), i.e. after every
    // thread has arrived for step t, which each does after its own gather of step t-1 -- the last
    // reader of that buffer.
    __shared__ uint64_t step_bar;
    if (tid == 0) mbar_init(&step_bar, THREADS / 32);   // one arrival per warp
    __syncthreads();
    if (cap_t == 0) capture();
    R J_ch[CPT], J_S[CPT], J_PLC[CPT];
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        J_ch[k] = J_S[k] = J_PLC[k] = (R)0;
        if (live[k]) {
            R rn;
            step_local<R>(q, c[k], p[k], s[k], r[k], vplc[k], J_ch[k], J_S[k], J_PLC[k], rn);
            r[k] = rn;      // r now holds r(t+1); r(t) was captured above if requested
        }
    }
    if ((tid & 31) == 0) mbar_arrive(&step_bar);   // state(0) was published before the __syncthreads above
    unsigned parity = 0;

    for (int t = 0; t < T - 1; ++t) {
        mbar_wait(&step_bar, parity);
        parity ^= 1u;
        // ---- Laplacian rows (csr_matvec order), all owned cells interleaved ---------------------
        R sc[CPT], sp[CPT];
#pragma unroll
        for (int k = 0; k < CPT; ++k) { sc[k] = (R)0; sp[k] = (R)0; }
#pragma unroll
        for (int e = 0; e < EMAX; ++e) {
#pragma unroll
            for (int k = 0; k < CPT; ++k) {
                if (e < wlen[k]) {   // warp-uniform
                    const uint32_t j = (nb[k][e >> 1] >> ((e & 1) * 16)) & 0xffffu;
                    const R2 v = cur[j];
                    if (UNIT) {
                        sc[k] = S::add(sc[k], v.x);
                        sp[k] = S::add(sp[k], v.y);
                    } else {
                        const R w = (j == ZERO_ENT) ? (R)0 : (R)__ldg(a.vals + rowbeg[k] + e);
                        sc[k] = S::add(sc[k], S::mul(w, v.x));
                        sp[k] = S::add(sp[k], S::mul(w, v.y));
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            if (live[k]) {
                const R lap_c = S::mul(q.D_c, sc[k]);                          // :346
                const R lap_p = S::mul(q.D_p, sp[k]);                          // :347
                R cn, pn;
                step_finish<R>(q, c[k], p[k], lap_c, lap_p, J_ch[k], J_S[k], J_PLC[k], cn, pn);
                c[k] = cn; p[k] = pn;
                const int slot = k * THREADS + tid;
                nxt[slot] = S::make2(cn, pn);
                if (COPIES == 2) nxt[np + (slot ^ 4)] = S::make2(cn, pn);
                if (UNIT) nxt[COPIES * np + slot] = S::make2(S::mul(dcoef[k], cn), S::mul(dcoef[k], pn));
            }
        }
        // one arrival per warp: __syncwarp orders the other lanes' stores before lane 0's release
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&step_bar);
        // ---- own-state work for the next step, overlapping the other warps' gathers --------------
#pragma unroll
        for (int k = 0; k < CPT; ++k)
            if (live[k]) s[k] = S::div(S::sub(q.c_tot, c[k]), q.beta);         // :95
        if (t + 1 == cap_t) capture();   // uniform across the CTA; r[] holds r(t+1) here
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            if (live[k]) {
                R rn;
                step_local<R>(q, c[k], p[k], s[k], r[k], vplc[k], J_ch[k], J_S[k], J_PLC[k], rn);
                r[k] = rn;
            }
        }
        const R2 *tmp = cur; cur = nxt; nxt = const_cast<R2 *>(tmp);
    }
}

}  // namespace cab200
