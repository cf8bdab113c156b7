/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product; nothing under calcium_b200/
 * may import, link or execute this file.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker / reported CPU baseline.
 *
 * Plain-C restatement of the reference's CPU algorithm for the simulation hot path:
 *   - explicit-Euler c/p/s/r integration: reference core/pouch.py:338-366 (time loop) and
 *     core/pouch.py:75-104 (_simulate_time_step), Laplacian products as scipy's csr_matvec does
 *     them at core/pouch.py:346-347 (per row: sum = 0; sum += A[jj] * x[col[jj]] in stored,
 *     i.e. ascending-column, order; scaled by D afterwards);
 *   - frame image / active mask / instance mask: core/pouch.py:418-427, 535-546, 563-565 and
 *     utils/sam2_data_generator.py:31-35.
 *
 * Arithmetic contract ("strict"): every operation is an individually rounded IEEE-754 binary64
 * operation in the source order of the reference's Python expressions; x**3 is (x*x)*x (what
 * numba emits for a constant integer power; the shipped reference is numba-jitted).  Build with
 * -ffp-contract=off and without -ffast-math so the compiler cannot fuse or reorder.
 *
 * Parity pin: checked against outputs of the reference itself (numba fastmath build and its
 * py_func) on committed fixtures, see oracle/make_golden.py and tests/test_oracle_golden.py.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { P_K1 = 0, P_KA, P_KP, P_K2, P_VSERCA, P_KSERCA, P_KPLC, P_K5, P_CTOT, P_BETA,
       P_KTAU, P_TAUMAX, P_KI, P_DC, P_DP, P_DT, P_COUNT };

static void spmv(int n, const int32_t *rowptr, const int32_t *col, const double *val,
                 const double *x, double *y)
{
    for (int i = 0; i < n; ++i) {
        double sum = 0.0;
        for (int jj = rowptr[i]; jj < rowptr[i + 1]; ++jj)
            sum += val[jj] * x[col[jj]];
        y[i] = sum;
    }
}

/* One pouch.  frames_out is [n_frames][4][n] (cell fastest).  Returns 0, or 1 if a non-finite
 * value appeared in a captured frame. */
int oracle_simulate(int n, const int32_t *rowptr, const int32_t *col, const double *val,
                    const double *prm, const double *c0, const double *p0, const double *r0,
                    const double *vplc, int T, int n_frames, const int32_t *frame_steps,
                    double *frames_out)
{
    const double k_1 = prm[P_K1], k_a = prm[P_KA], k_p = prm[P_KP], k_2 = prm[P_K2];
    const double V_S = prm[P_VSERCA], K_S = prm[P_KSERCA], K_PLC = prm[P_KPLC], K_5 = prm[P_K5];
    const double c_tot = prm[P_CTOT], beta = prm[P_BETA], k_tau = prm[P_KTAU];
    const double tau_max = prm[P_TAUMAX], k_i = prm[P_KI], D_c = prm[P_DC], D_p = prm[P_DP];
    const double dt = prm[P_DT];

    double *buf = (double *)malloc(sizeof(double) * (size_t)n * 10);
    if (!buf) return -1;
    double *c = buf, *p = buf + n, *s = buf + 2 * n, *r = buf + 3 * n;
    double *cn = buf + 4 * n, *pn = buf + 5 * n, *sn = buf + 6 * n, *rn = buf + 7 * n;
    double *lc = buf + 8 * n, *lp = buf + 9 * n;
    int status = 0, next = 0;

    for (int i = 0; i < n; ++i) {
        c[i] = c0[i];
        p[i] = p0[i];
        s[i] = (c_tot - c0[i]) / beta;          /* core/pouch.py:322 */
        r[i] = r0[i];                            /* core/pouch.py:325 */
    }
    const double KS2 = K_S * K_S, KP2 = K_PLC * K_PLC;
    const double k_tau_4 = k_tau * k_tau * k_tau * k_tau;   /* ((k*k)*k)*k, core/pouch.py:99 */
    const double rec_den = tau_max * k_tau_4;

    for (int step = 0; step < T; ++step) {
        if (next < n_frames && frame_steps[next] == step) {
            double *f = frames_out + (size_t)next * 4 * n;
            memcpy(f, c, sizeof(double) * n);
            memcpy(f + n, p, sizeof(double) * n);
            memcpy(f + 2 * n, s, sizeof(double) * n);
            memcpy(f + 3 * n, r, sizeof(double) * n);
            for (int i = 0; i < 4 * n; ++i)
                if (!isfinite(f[i])) status = 1;
            ++next;
        }
        if (step == T - 1) break;
        spmv(n, rowptr, col, val, c, lc);
        spmv(n, rowptr, col, val, p, lp);
        for (int i = 0; i < n; ++i) {
            const double ca = c[i], ipt = p[i], sv = s[i], rv = r[i];
            const double lap_c = D_c * lc[i];               /* core/pouch.py:346 */
            const double lap_p = D_p * lp[i];               /* core/pouch.py:347 */
            const double num = (rv * ca) * ipt;              /* :77 */
            const double d1 = k_a + ca, d2 = k_p + ipt;      /* :78-79 */
            const double x = num / (d1 * d2);                /* :80 */
            const double x3 = (x * x) * x;
            const double J_ch = (k_1 * x3 + k_2) * (sv - ca);      /* :81 */
            const double ca_sq = ca * ca;                          /* :84 */
            const double J_S = (V_S * ca_sq) / (ca_sq + KS2);      /* :85 */
            const double c_new = ca + dt * ((lap_c + J_ch) - J_S); /* :87 */
            const double J_PLC = (vplc[i] * ca_sq) / (ca_sq + KP2);        /* :91 */
            const double p_new = ipt + dt * ((lap_p + J_PLC) - K_5 * ipt); /* :92 */
            const double s_new = (c_tot - c_new) / beta;                   /* :95 */
            const double ca_4 = ca_sq * ca_sq;                             /* :98 */
            const double rec = (k_tau_4 + ca_4) / rec_den;                 /* :100 */
            const double inact = 1.0 - (rv * (k_i + ca)) / k_i;            /* :101 */
            const double r_new = rv + (dt * rec) * inact;                  /* :102 */
            cn[i] = c_new; pn[i] = p_new; sn[i] = s_new; rn[i] = r_new;
        }
        double *t;
        t = c; c = cn; cn = t;
        t = p; p = pn; pn = t;
        t = s; s = sn; sn = t;
        t = r; r = rn; rn = t;
    }
    free(buf);
    return status;
}

/* float32 variant of the same sequence (numpy float32 evaluation of the reference formulas);
 * used to report the FP32 mode's error against an identically-rounded CPU run. */
int oracle_simulate_f32(int n, const int32_t *rowptr, const int32_t *col, const double *val,
                        const double *prm, const double *c0, const double *p0, const double *r0,
                        const double *vplc, int T, int n_frames, const int32_t *frame_steps,
                        float *frames_out)
{
    const float k_1 = (float)prm[P_K1], k_a = (float)prm[P_KA], k_p = (float)prm[P_KP];
    const float k_2 = (float)prm[P_K2], V_S = (float)prm[P_VSERCA], K_S = (float)prm[P_KSERCA];
    const float K_PLC = (float)prm[P_KPLC], K_5 = (float)prm[P_K5], c_tot = (float)prm[P_CTOT];
    const float beta = (float)prm[P_BETA], k_tau = (float)prm[P_KTAU];
    const float tau_max = (float)prm[P_TAUMAX], k_i = (float)prm[P_KI];
    const float D_c = (float)prm[P_DC], D_p = (float)prm[P_DP], dt = (float)prm[P_DT];
    float *buf = (float *)malloc(sizeof(float) * (size_t)n * 9);
    if (!buf) return -1;
    float *c = buf, *p = buf + n, *s = buf + 2 * n, *r = buf + 3 * n;
    float *cn = buf + 4 * n, *pn = buf + 5 * n, *sn = buf + 6 * n, *rn = buf + 7 * n;
    float *v = buf + 8 * n;
    int status = 0, next = 0;
    for (int i = 0; i < n; ++i) {
        c[i] = (float)c0[i]; p[i] = (float)p0[i]; r[i] = (float)r0[i]; v[i] = (float)vplc[i];
        s[i] = (c_tot - c[i]) / beta;
    }
    const float KS2 = K_S * K_S, KP2 = K_PLC * K_PLC;
    const float k_tau_4 = k_tau * k_tau * k_tau * k_tau;
    const float rec_den = tau_max * k_tau_4;
    for (int step = 0; step < T; ++step) {
        if (next < n_frames && frame_steps[next] == step) {
            float *f = frames_out + (size_t)next * 4 * n;
            memcpy(f, c, sizeof(float) * n); memcpy(f + n, p, sizeof(float) * n);
            memcpy(f + 2 * n, s, sizeof(float) * n); memcpy(f + 3 * n, r, sizeof(float) * n);
            for (int i = 0; i < 4 * n; ++i) if (!isfinite(f[i])) status = 1;
            ++next;
        }
        if (step == T - 1) break;
        for (int i = 0; i < n; ++i) {
            float sc = 0.0f, sp = 0.0f;
            for (int jj = rowptr[i]; jj < rowptr[i + 1]; ++jj) {
                const float w = (float)val[jj];
                sc += w * c[col[jj]];
                sp += w * p[col[jj]];
            }
            const float ca = c[i], ipt = p[i], sv = s[i], rv = r[i];
            const float lap_c = D_c * sc, lap_p = D_p * sp;
            const float x = ((rv * ca) * ipt) / ((k_a + ca) * (k_p + ipt));
            const float x3 = (x * x) * x;
            const float J_ch = (k_1 * x3 + k_2) * (sv - ca);
            const float ca_sq = ca * ca;
            const float J_S = (V_S * ca_sq) / (ca_sq + KS2);
            const float c_new = ca + dt * ((lap_c + J_ch) - J_S);
            const float J_PLC = (v[i] * ca_sq) / (ca_sq + KP2);
            const float p_new = ipt + dt * ((lap_p + J_PLC) - K_5 * ipt);
            const float s_new = (c_tot - c_new) / beta;
            const float ca_4 = ca_sq * ca_sq;
            const float rec = (k_tau_4 + ca_4) / rec_den;
            const float inact = 1.0f - (rv * (k_i + ca)) / k_i;
            const float r_new = rv + (dt * rec) * inact;
            cn[i] = c_new; pn[i] = p_new; sn[i] = s_new; rn[i] = r_new;
        }
        float *t;
        t = c; c = cn; cn = t;  t = p; p = pn; pn = t;
        t = s; s = sn; sn = t;  t = r; r = rn; rn = t;
    }
    free(buf);
    return status;
}

/* ---- ensemble: pouches are independent (reference main.py:202 runs them one after another);
 * here they are spread over host threads so the CPU baseline can use every core. ------------- */
typedef struct {
    int n_pouch, next;
    pthread_mutex_t mu;
    const int32_t *geom_id, *g_ncell;
    const int64_t *g_rowptr_off, *g_nnz_off;
    const int32_t *rowptr, *col;
    const double *val;
    const int64_t *cell_off;
    const double *params, *c0, *p0, *r0, *vplc;
    int T, n_frames;
    const int32_t *frame_steps;
    double *frames_out;
    int32_t *status_out;
} batch_t;

static void *worker(void *arg)
{
    batch_t *b = (batch_t *)arg;
    for (;;) {
        pthread_mutex_lock(&b->mu);
        int i = b->next++;
        pthread_mutex_unlock(&b->mu);
        if (i >= b->n_pouch) break;
        const int g = b->geom_id[i];
        const int n = b->g_ncell[g];
        const int64_t co = b->cell_off[i];
        b->status_out[i] = oracle_simulate(
            n, b->rowptr + b->g_rowptr_off[g], b->col + b->g_nnz_off[g], b->val + b->g_nnz_off[g],
            b->params + (size_t)i * P_COUNT, b->c0 + co, b->p0 + co, b->r0 + co, b->vplc + co,
            b->T, b->n_frames, b->frame_steps, b->frames_out + (size_t)co * 4 * b->n_frames);
    }
    return NULL;
}

/* Same argument meaning as cab200_sim_batch (include/cab200.h), host pointers. */
int oracle_sim_batch(int n_pouch, const int32_t *geom_id, int n_geom, const int32_t *g_ncell,
                     const int64_t *g_rowptr_off, const int64_t *g_nnz_off,
                     const int32_t *rowptr, const int32_t *col, const double *val,
                     const int64_t *cell_off, const double *params,
                     const double *c0, const double *p0, const double *r0, const double *vplc,
                     int T, int n_frames, const int32_t *frame_steps,
                     double *frames_out, int32_t *status_out, int n_threads)
{
    (void)n_geom;
    batch_t b = { n_pouch, 0, PTHREAD_MUTEX_INITIALIZER, geom_id, g_ncell, g_rowptr_off, g_nnz_off,
                  rowptr, col, val, cell_off, params, c0, p0, r0, vplc, T, n_frames, frame_steps,
                  frames_out, status_out };
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_pouch) n_threads = n_pouch > 0 ? n_pouch : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, worker, &b);
    for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    free(th);
    return 0;
}

/* ---- per-frame raster (gather through the two static label maps) ------------------------- */

/* image (H,W,2) u8, active mask (H,W) i32, instance mask (H,W) u16 for one frame.
 * label_img: cv2 paint-order map (0 = background, id = cell+1); label_mask: cells_mask.
 * Returns the number of active cells.  quirk != 0 reproduces the reference's off-by-one between
 * get_cell_masks' 1-based ids and get_active_cells' 0-based ids
 * (utils/sam2_data_generator.py:34-35 vs core/pouch.py:544,564). */
int oracle_raster_frame(int n, const double *ca, const uint16_t *label_img,
                        const uint16_t *label_mask, int H, int W, double threshold, int quirk,
                        uint8_t *image_out, int32_t *active_out, uint16_t *instance_out)
{
    double mx = ca[0];
    for (int i = 1; i < n; ++i) if (ca[i] > mx) mx = ca[i];      /* np.max */
    const double vmax = (0.5 > mx) ? 0.5 : mx;                    /* max(np.max(..), 0.5), :422 */
    uint8_t *gray = (uint8_t *)malloc((size_t)n + 1);
    int32_t *rank1 = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n + 2));
    uint8_t *act = (uint8_t *)malloc((size_t)n + 2);
    int n_act = 0;
    for (int i = 0; i < n; ++i) {
        double nv = (ca[i] - 0.0) / (vmax - 0.0);                 /* :426 */
        if (nv < 0.0) nv = 0.0; else if (nv > 1.0) nv = 1.0;
        gray[i] = (uint8_t)(int)(255.0 * nv);                      /* :427 */
        act[i] = ca[i] > threshold;                                /* :564 */
        rank1[i] = act[i] ? ++n_act : 0;
    }
    act[n] = 0; rank1[n] = 0;
    for (size_t px = 0; px < (size_t)H * W; ++px) {
        const int li = label_img[px];
        if (image_out) {
            image_out[2 * px] = li ? gray[li - 1] : 0;
            image_out[2 * px + 1] = li ? 255 : 0;
        }
        const int lm = label_mask[px];
        const int a = (lm && act[lm - 1]) ? lm : 0;               /* :544 */
        if (active_out) active_out[px] = a;
        if (instance_out) {
            if (quirk) instance_out[px] = (uint16_t)((a <= n && act[a]) ? rank1[a] : 0);
            else       instance_out[px] = (uint16_t)(a ? rank1[a - 1] : 0);
        }
    }
    free(gray); free(rank1); free(act);
    return n_act;
}
