// integrate.cu -- host launcher of K1 (see integrate.cuh) behind cab200_sim_batch.
#include <algorithm>
#include <atomic>
#include <mutex>
#include <vector>

#include "integrate_pick.cuh"

namespace cab200 {

struct Shape { int n_max, threads, cpt, minb; };

// Default planner: CTA shape by cell count.  One CTA per pouch; small pouches get several CTAs per
// SM (MINB) because one step is latency bound (~6 dependent FP64 divides + barrier).
static const Shape kShapes[] = {
    {128, 128, 1, 8},
    {256, 256, 1, 4},
    {512, 256, 2, 2},
    {1024, 512, 2, 1},
    {1536, 512, 3, 1},
    {2048, 1024, 2, 1},
    {3072, 1024, 3, 1},
    {4096, 1024, 4, 1},
};

void default_shape(int n, int *threads, int *cpt)
{
    *threads = 0; *cpt = 0;
    for (const Shape &s : kShapes)
        if (n <= s.n_max) { *threads = s.threads; *cpt = s.cpt; return; }
}

// Two value copies (conflict dodging, see integrate.cuh) whenever they fit in shared memory.
int sim_copies(int n, bool unit, size_t elem2, int nslot)
{
    if (!unit) return 1;
    const int np = (n + 31) & ~31;
    return integrate_smem_bytes(np, nslot, elem2, 2, true) <= 227 * 1024 ? 2 : 1;
}

// frames_tmp (slot order, see integrate.cuh) -> frames_out (original cell order).  One CTA per
// (frame, pouch of the launch): gathers inside one 4*NS-element block (L1/L2 resident), writes coalesced.
template <typename R>
__global__ void __launch_bounds__(256) unpermute_kernel(const R *tmp, R *out, const int32_t *pouch_list,
                                                        const int64_t *cell_off, const uint16_t *plan, int n, int ncta,
                                                        int nslot, int n_frames)
{
    const int q = blockIdx.y, fi = blockIdx.x;
    const int64_t coff = cell_off[pouch_list[q]];
    const size_t ns = (size_t)ncta * nslot;
    const R *src = tmp + ((size_t)q * n_frames + fi) * 4 * ns;
    R *dst = out + (size_t)coff * 4 * n_frames + (size_t)fi * 4 * n;
    for (int cell = threadIdx.x; cell < n; cell += blockDim.x) {
        // cluster plan: part_of[n] | slot_of[n] | ...;  single-CTA plan: slot_of[n] | ...
        const size_t g = (ncta > 1) ? (size_t)plan[cell] * nslot + plan[n + cell] : (size_t)plan[cell];
#pragma unroll
        for (int st = 0; st < 4; ++st) dst[(size_t)st * n + cell] = src[st * ns + g];
    }
}

static std::mutex g_plan_mu;
static int g_force_nmax = 0, g_force_threads = 0, g_force_cpt = 0;

// Side streams for launches of different geometries inside one cab200_sim_batch call.  A mixed-size ensemble is
// one persistent kernel launch per geometry; on the caller's stream alone those run back to back, although a small
// group leaves most SMs idle (8 GPUs x 32 mixed pouches: 52.7 ms as a sequence, the longest group alone ~20 ms).
// Group 0 stays on the caller's stream, group k > 0 goes to side stream (k-1) mod 3 of the current device, forked
// and joined with events, so the call remains stream-ordered for the caller.  A few sets per device: host threads
// that drive the same device concurrently (devices=[0, 0]) must not share events.
constexpr int CAB_SIDE_STREAMS = 3, CAB_SIDE_SETS = 4, CAB_MAX_DEVICES = 64;
struct SideStreams {
    std::atomic<int> busy{0};
    bool ready = false;
    cudaStream_t s[CAB_SIDE_STREAMS] = {};
    cudaEvent_t fork = nullptr, join[CAB_SIDE_STREAMS] = {};
};
static SideStreams g_side[CAB_MAX_DEVICES][CAB_SIDE_SETS];

static SideStreams *acquire_side_streams(int device)
{
    if (device < 0 || device >= CAB_MAX_DEVICES) return nullptr;
    for (int k = 0; k < CAB_SIDE_SETS; ++k) {
        SideStreams &ss = g_side[device][k];
        int expect = 0;
        if (!ss.busy.compare_exchange_strong(expect, 1)) continue;
        if (!ss.ready) {
            bool ok = cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming) == cudaSuccess;
            for (int i = 0; ok && i < CAB_SIDE_STREAMS; ++i)
                ok = cudaStreamCreateWithFlags(&ss.s[i], cudaStreamNonBlocking) == cudaSuccess &&
                     cudaEventCreateWithFlags(&ss.join[i], cudaEventDisableTiming) == cudaSuccess;
            if (!ok) { cudaGetLastError(); ss.busy = 0; return nullptr; }
            ss.ready = true;
        }
        return &ss;
    }
    return nullptr;                      // every set is in use by another host thread: stay on the caller's stream
}

// Hands out the stream of the k-th non-empty geometry group and joins the side streams back into the caller's
// stream when it goes out of scope (error returns included), so the call stays stream-ordered.
struct GroupStreams {
    cudaStream_t main;
    SideStreams *ss = nullptr;
    bool used[CAB_SIDE_STREAMS] = {};
    int next = 0;
    GroupStreams(cudaStream_t main_, int n_groups) : main(main_)
    {
        if (n_groups < 2) return;
        int dev = -1;
        if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return; }
        ss = acquire_side_streams(dev);
        if (ss && cudaEventRecord(ss->fork, main) != cudaSuccess) { cudaGetLastError(); ss->busy = 0; ss = nullptr; }
    }
    bool concurrent() const { return ss != nullptr; }
    cudaStream_t take()
    {
        const int k = next++;
        if (!ss || k == 0) return main;
        const int i = (k - 1) % CAB_SIDE_STREAMS;
        if (!used[i]) {
            if (cudaStreamWaitEvent(ss->s[i], ss->fork, 0) != cudaSuccess) { cudaGetLastError(); return main; }
            used[i] = true;
        }
        return ss->s[i];
    }
    ~GroupStreams()
    {
        if (!ss) return;
        for (int i = 0; i < CAB_SIDE_STREAMS; ++i)
            if (used[i] && cudaEventRecord(ss->join[i], ss->s[i]) == cudaSuccess) cudaStreamWaitEvent(main, ss->join[i], 0);
        ss->busy = 0;
    }
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace cab200

using namespace cab200;

extern "C" int cab200_sim_set_plan(int n_max, int threads, int cells_per_thread)
{
    std::lock_guard<std::mutex> lk(g_plan_mu);
    if (threads == 0) { g_force_nmax = g_force_threads = g_force_cpt = 0; return CAB200_OK; }
    if (!pick_kernel_f64(threads, cells_per_thread, 5, true, 1)) {
        set_error("cab200_sim_set_plan: unsupported shape threads=%d cells_per_thread=%d", threads,
                  cells_per_thread);
        return CAB200_EINVAL;
    }
    g_force_nmax = n_max; g_force_threads = threads; g_force_cpt = cells_per_thread;
    return CAB200_OK;
}

extern "C" int cab200_sim_layout(int n, int unit, int precision, int64_t out[4])
{
    if (!out || n <= 0 || (precision != CAB200_FP64_STRICT && precision != CAB200_FP32)) {
        set_error("cab200_sim_layout: bad argument");
        return CAB200_EINVAL;
    }
    int threads = 0, cpt = 0;
    {
        std::lock_guard<std::mutex> lk(g_plan_mu);
        if (g_force_threads && n <= g_force_nmax && g_force_threads * g_force_cpt >= n) {
            threads = g_force_threads; cpt = g_force_cpt;
        }
    }
    if (!threads) default_shape(n, &threads, &cpt);
    if (!threads) { set_error("cab200_sim_layout: %d cells exceed one CTA (4096)", n); return CAB200_ESIZE; }
    const size_t elem2 = (precision == CAB200_FP64_STRICT) ? 16 : 8;
    const int copies = sim_copies(n, unit != 0, elem2, threads * cpt);
    out[0] = threads; out[1] = cpt; out[2] = copies;
    out[3] = (int64_t)integrate_smem_bytes((n + 31) & ~31, threads * cpt, elem2, copies, unit != 0);
    return CAB200_OK;
}

extern "C" int cab200_cluster_layout(int n, int n_cta, int precision, int64_t out[4])
{
    if (!out || n <= 0 || (n_cta != 2 && n_cta != 3 && n_cta != 4 && n_cta != 8) ||
        (precision != CAB200_FP64_STRICT && precision != CAB200_FP32)) {
        set_error("cab200_cluster_layout: bad argument (cluster sizes 2, 3, 4 and 8 are supported)");
        return CAB200_EINVAL;
    }
    int threads = 0, cpt = 0;
    cluster_shape((n + n_cta - 1) / n_cta, &threads, &cpt);
    if (n_cta == 8 && threads && threads < 416) { threads = 416; cpt = 2; }      // clusters of 8: 416- / 512-thread shapes only
    if (n_cta == 3 && threads) { threads = 512; cpt = 3; }                       // clusters of 3: the 512 x 3 shape only
    if (!threads) { set_error("cab200_cluster_layout: %d cells exceed %d CTAs x 1536", n, n_cta); return CAB200_ESIZE; }
    const size_t elem2 = (precision == CAB200_FP64_STRICT) ? 16 : 8;
    const int copies = integrate_smem_bytes(0, threads * cpt, elem2, 2, true, CAB_HMAX) <= 227 * 1024 ? 2 : 1;
    out[0] = threads; out[1] = cpt; out[2] = copies;
    out[3] = (int64_t)integrate_smem_bytes(0, threads * cpt, elem2, copies, true, CAB_HMAX, n);
    return CAB200_OK;
}

extern "C" int cab200_cluster_capacity(int n, int max_row, int n_cta, int precision)
{
    int64_t lay[4];
    if (int rc = cab200_cluster_layout(n, n_cta, precision, lay)) return rc;
    if (max_row < 0 || max_row > 16) { set_error("cab200_cluster_capacity: longest CSR row %d outside [0,16]", max_row); return CAB200_ESIZE; }
    const int ep = k1_ep(max_row);
    KernelFn fn = (precision == CAB200_FP64_STRICT) ? pick_cluster_kernel_f64((int)lay[0], (int)lay[1], n_cta, ep, (int)lay[2])
                                                    : pick_cluster_kernel_f32((int)lay[0], (int)lay[1], n_cta, ep, (int)lay[2]);
    if (!fn) { set_error("cab200_cluster_capacity: no cluster kernel for %dx%d x%d", (int)lay[0], (int)lay[1], n_cta); return CAB200_EINVAL; }
    if (cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay[3]) != cudaSuccess) {
        set_error("cab200_cluster_capacity: %s", cudaGetErrorString(cudaGetLastError()));
        return CAB200_ECUDA;
    }
    int n_sm = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_cta * std::max(1, n_sm / n_cta)), 1, 1);
    cfg.blockDim = dim3((unsigned)lay[0], 1, 1);
    cfg.dynamicSmemBytes = (size_t)lay[3];
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)n_cta; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&n_clusters, (const void *)fn, &cfg) != cudaSuccess) {
        set_error("cab200_cluster_capacity: %s", cudaGetErrorString(cudaGetLastError()));
        return CAB200_ECUDA;
    }
    return n_clusters;
}

extern "C" size_t cab200_sim_frames_tmp_bytes(int n_cells, int n_cta, int unit, int precision, int n_pouch, int n_frames)
{
    int64_t lay[4];
    const int rc = (n_cta > 1) ? cab200_cluster_layout(n_cells, n_cta, precision, lay) : cab200_sim_layout(n_cells, unit, precision, lay);
    if (rc != CAB200_OK || n_pouch <= 0 || n_frames <= 0) return 0;
    const size_t ns = (size_t)(n_cta > 1 ? n_cta : 1) * (size_t)lay[0] * (size_t)lay[1];
    return (size_t)n_pouch * n_frames * 4 * ns * (precision == CAB200_FP64_STRICT ? 8 : 4);
}

extern "C" size_t cab200_sim_scratch_bytes(int n_pouch, int n_frames)
{
    // pouch_list [P] i32 + cell_off [P+1] i64 + frame_steps [F] i32, each 16-byte aligned
    return align_up(sizeof(int32_t) * (size_t)std::max(n_pouch, 1), 16) +
           align_up(sizeof(int64_t) * ((size_t)n_pouch + 1), 16) +
           align_up(sizeof(int32_t) * (size_t)std::max(n_frames, 1), 16);
}

extern "C" int cab200_sim_batch(
    int n_pouch, const int32_t *geom_id, int n_geom, const int32_t *g_ncell,
    const int32_t *g_max_row, const int32_t *g_unit, const int64_t *g_rowptr_off,
    const int64_t *g_nnz_off, const int32_t *rowptr, const int32_t *colidx, const double *vals,
    const int32_t *g_cluster, const int64_t *g_plan_off, const uint16_t *slot_plan, const int64_t *cell_off, const double *params, const double *c0, const double *p0,
    const double *r0, const double *vplc, int T, int n_frames, const int32_t *frame_steps,
    void *frames_out, int32_t *status_out, int precision, void *scratch, size_t scratch_bytes,
    void *frames_tmp, size_t frames_tmp_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_pouch < 0 || n_geom <= 0 || T < 1 || n_frames < 0) {
        set_error("cab200_sim_batch: bad sizes (n_pouch=%d n_geom=%d T=%d n_frames=%d)", n_pouch,
                  n_geom, T, n_frames);
        return CAB200_EINVAL;
    }
    if (precision != CAB200_FP64_STRICT && precision != CAB200_FP32) {
        set_error("cab200_sim_batch: precision must be 64 or 32, got %d", precision);
        return CAB200_EINVAL;
    }
    if (n_pouch == 0) return CAB200_OK;
    if (!geom_id || !g_ncell || !g_max_row || !g_unit || !g_rowptr_off || !g_nnz_off || !rowptr ||
        !colidx || !vals || !cell_off || !params || !c0 || !p0 || !r0 || !vplc || !status_out ||
        (n_frames > 0 && (!frame_steps || !frames_out)) || !scratch) {
        set_error("cab200_sim_batch: null pointer argument");
        return CAB200_EINVAL;
    }
    if (scratch_bytes < cab200_sim_scratch_bytes(n_pouch, n_frames)) {
        set_error("cab200_sim_batch: scratch too small (%zu < %zu)", scratch_bytes,
                  cab200_sim_scratch_bytes(n_pouch, n_frames));
        return CAB200_ESIZE;
    }
    for (int f = 0; f < n_frames; ++f) {
        if (frame_steps[f] < 0 || frame_steps[f] > T - 1 || (f > 0 && frame_steps[f] <= frame_steps[f - 1])) {
            set_error("cab200_sim_batch: frame_steps must be strictly ascending within [0, T-1]");
            return CAB200_EINVAL;
        }
    }
    for (int p = 0; p < n_pouch; ++p) {
        const int g = geom_id[p];
        if (g < 0 || g >= n_geom || cell_off[p + 1] - cell_off[p] != g_ncell[g]) {
            set_error("cab200_sim_batch: pouch %d: geom_id/cell_off inconsistent", p);
            return CAB200_EINVAL;
        }
    }

    // ---- device copies of the small metadata (scratch layout as in cab200_sim_scratch_bytes) ----
    unsigned char *sp = (unsigned char *)scratch;
    int32_t *d_list = (int32_t *)sp;
    sp += align_up(sizeof(int32_t) * (size_t)n_pouch, 16);
    int64_t *d_cell_off = (int64_t *)sp;
    sp += align_up(sizeof(int64_t) * ((size_t)n_pouch + 1), 16);
    int32_t *d_frame_steps = (int32_t *)sp;

    // pouch ids grouped by geometry
    std::vector<int32_t> list;
    list.reserve(n_pouch);
    std::vector<int> g_begin(n_geom + 1, 0);
    for (int g = 0; g < n_geom; ++g) {
        g_begin[g] = (int)list.size();
        for (int p = 0; p < n_pouch; ++p)
            if (geom_id[p] == g) list.push_back(p);
    }
    g_begin[n_geom] = (int)list.size();
    CAB_CUDA(cudaMemcpyAsync(d_list, list.data(), sizeof(int32_t) * list.size(),
                             cudaMemcpyHostToDevice, stream));
    CAB_CUDA(cudaMemcpyAsync(d_cell_off, cell_off, sizeof(int64_t) * ((size_t)n_pouch + 1),
                             cudaMemcpyHostToDevice, stream));
    if (n_frames > 0)
        CAB_CUDA(cudaMemcpyAsync(d_frame_steps, frame_steps, sizeof(int32_t) * (size_t)n_frames,
                                 cudaMemcpyHostToDevice, stream));
    CAB_CUDA(cudaMemsetAsync(status_out, 0, sizeof(int32_t) * (size_t)n_pouch, stream));
    // the host vectors above are pageable: the copies are staged before cudaMemcpyAsync returns.

    int force_nmax, force_threads, force_cpt;
    {
        std::lock_guard<std::mutex> lk(g_plan_mu);
        force_nmax = g_force_nmax; force_threads = g_force_threads; force_cpt = g_force_cpt;
    }

    int n_groups = 0;
    for (int g = 0; g < n_geom; ++g) n_groups += g_begin[g + 1] > g_begin[g];
    GroupStreams gs(stream, n_groups);          // forks after the metadata copies above, joins on scope exit
    size_t tmp_cursor = 0;                      // concurrent groups take disjoint pieces of frames_tmp

    // largest pouches first: their CTAs run longest, so they should not start behind the short ones
    std::vector<int> g_order(n_geom);
    for (int g = 0; g < n_geom; ++g) g_order[g] = g;
    std::stable_sort(g_order.begin(), g_order.end(), [&](int a_, int b_) { return g_ncell[a_] > g_ncell[b_]; });

    for (int gi = 0; gi < n_geom; ++gi) {
        const int g = g_order[gi];
        const int count = g_begin[g + 1] - g_begin[g];
        if (count == 0) continue;
        const int n = g_ncell[g];
        if (n <= 0) { set_error("cab200_sim_batch: geometry %d has no cells", g); return CAB200_EINVAL; }
        stream = gs.take();
        unsigned char *const tmp_base = frames_tmp ? (unsigned char *)frames_tmp + tmp_cursor : nullptr;
        const size_t tmp_avail = frames_tmp ? frames_tmp_bytes - tmp_cursor : 0;
        const int ncta = g_cluster ? g_cluster[g] : 1;
        const int64_t plan_off = (slot_plan && g_plan_off) ? g_plan_off[g] : -1;
        if (ncta > 1) {
            // ---- pouch split over a thread-block cluster ---------------------------------------
            if (ncta != 2 && ncta != 3 && ncta != 4 && ncta != 8) { set_error("cab200_sim_batch: geometry %d: cluster size %d (2, 3, 4 or 8)", g, ncta); return CAB200_EINVAL; }
            if (plan_off < 0) { set_error("cab200_sim_batch: geometry %d: a cluster needs a cab200_plan_cluster plan", g); return CAB200_EINVAL; }
            if (!g_unit[g]) { set_error("cab200_sim_batch: geometry %d: clusters need the unit-weight Laplacian", g); return CAB200_EINVAL; }
            const int max_row_c = g_max_row[g];
            if (max_row_c > 16 || max_row_c < 0) { set_error("cab200_sim_batch: geometry %d: longest CSR row %d outside [0,16]", g, max_row_c); return CAB200_ESIZE; }
            int cth = 0, ccpt = 0;
            cluster_shape((n + ncta - 1) / ncta, &cth, &ccpt);
            if (ncta == 8 && cth && cth < 416) { cth = 416; ccpt = 2; }
            if (ncta == 3 && cth) { cth = 512; ccpt = 3; }
            if (!cth) { set_error("cab200_sim_batch: geometry %d: %d cells exceed %d CTAs x 1536", g, n, ncta); return CAB200_ESIZE; }
            const size_t el2 = (precision == CAB200_FP64_STRICT) ? 16 : 8;
            const int ccopies = integrate_smem_bytes(0, cth * ccpt, el2, 2, true, CAB_HMAX) <= 227 * 1024 ? 2 : 1;
            const int cep = k1_ep(max_row_c);
            KernelFn cfn = (precision == CAB200_FP64_STRICT) ? pick_cluster_kernel_f64(cth, ccpt, ncta, cep, ccopies)
                                                             : pick_cluster_kernel_f32(cth, ccpt, ncta, cep, ccopies);
            if (!cfn) { set_error("cab200_sim_batch: no cluster kernel for %dx%d x%d", cth, ccpt, ncta); return CAB200_EINVAL; }
            const size_t csmem = integrate_smem_bytes(0, cth * ccpt, el2, ccopies, true, CAB_HMAX, n);
            CAB_CUDA(cudaFuncSetAttribute((const void *)cfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
            SimArgs ca;
            ca.pouch_list = d_list + g_begin[g];
            ca.cell_off = d_cell_off;
            ca.rowptr = rowptr + g_rowptr_off[g];
            ca.colidx = colidx + g_nnz_off[g];
            ca.vals = vals + g_nnz_off[g];
            ca.slot_plan = slot_plan + plan_off;
            ca.n = n; ca.np = (n + 31) & ~31;
            ca.params = params; ca.c0 = c0; ca.p0 = p0; ca.r0 = r0; ca.vplc = vplc;
            ca.T = T; ca.n_frames = n_frames; ca.frame_steps = d_frame_steps;
            ca.frames_out = frames_out; ca.status_out = status_out;
            const size_t tmp_need = (size_t)count * n_frames * 4 * ((size_t)ncta * cth * ccpt) * (el2 / 2);
            ca.frames_tmp = (tmp_base && n_frames > 0 && tmp_need <= tmp_avail && count <= 65535) ? tmp_base : nullptr;
            if (ca.frames_tmp && gs.concurrent()) tmp_cursor += align_up(tmp_need, 256);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(count * ncta), 1, 1);
            cfg.blockDim = dim3((unsigned)cth, 1, 1);
            cfg.dynamicSmemBytes = csmem;
            cfg.stream = stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = (unsigned)ncta; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            CAB_CUDA(cudaLaunchKernelEx(&cfg, cfn, ca));
            if (ca.frames_tmp) {
                const dim3 ug((unsigned)n_frames, (unsigned)count);
                if (precision == CAB200_FP64_STRICT)
                    unpermute_kernel<double><<<ug, 256, 0, stream>>>((const double *)ca.frames_tmp, (double *)frames_out, ca.pouch_list,
                                                                     d_cell_off, ca.slot_plan, n, ncta, cth * ccpt, n_frames);
                else
                    unpermute_kernel<float><<<ug, 256, 0, stream>>>((const float *)ca.frames_tmp, (float *)frames_out, ca.pouch_list,
                                                                    d_cell_off, ca.slot_plan, n, ncta, cth * ccpt, n_frames);
                CAB_CUDA(cudaGetLastError());
            }
            continue;
        }
        int threads = 0, cpt = 0;
        if (force_threads && n <= force_nmax && force_threads * force_cpt >= n) {
            threads = force_threads; cpt = force_cpt;
        } else {
            default_shape(n, &threads, &cpt);
        }
        if (!threads) {
            set_error("cab200_sim_batch: geometry %d has %d cells; one CTA holds at most 4096", g, n);
            return CAB200_ESIZE;
        }
        const int max_row = g_max_row[g];
        if (max_row > 16 || max_row < 0) {
            set_error("cab200_sim_batch: geometry %d: longest CSR row %d outside [0,16]", g, max_row);
            return CAB200_ESIZE;
        }
        const int ep = k1_ep(max_row);      // rows of up to 10 entries use the 5-pair variant
        // the unit-weight kernel needs the planner's cell -> slot map (rows with many entries on one side of their
        // diagonal must share a warp, k1_layout.cuh); without a plan the general-weights kernel runs: same bits, slower
        const bool unit = g_unit[g] != 0 && plan_off >= 0;
        const size_t elem2 = (precision == CAB200_FP64_STRICT) ? 16 : 8;
        const int np = (n + 31) & ~31;
        const int copies = sim_copies(n, unit, elem2, threads * cpt);
        KernelFn fn = (precision == CAB200_FP64_STRICT) ? pick_kernel_f64(threads, cpt, ep, unit, copies)
                                                        : pick_kernel_f32(threads, cpt, ep, unit, copies);
        if (!fn) { set_error("cab200_sim_batch: no kernel for shape %dx%d", threads, cpt); return CAB200_EINVAL; }
        const size_t smem = integrate_smem_bytes(np, threads * cpt, elem2, copies, unit);
        if (smem > 227 * 1024) {
            set_error("cab200_sim_batch: geometry %d needs %zu bytes of shared memory (> 227 KB)", g, smem);
            return CAB200_ESIZE;
        }
        CAB_CUDA(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

        SimArgs a;
        a.pouch_list = d_list + g_begin[g];
        a.cell_off = d_cell_off;
        a.rowptr = rowptr + g_rowptr_off[g];
        a.colidx = colidx + g_nnz_off[g];
        a.vals = vals + g_nnz_off[g];
        a.n = n;
        a.slot_plan = (plan_off >= 0) ? slot_plan + plan_off : nullptr;
        a.np = np;
        a.params = params;
        a.c0 = c0; a.p0 = p0; a.r0 = r0; a.vplc = vplc;
        a.T = T; a.n_frames = n_frames;
        a.frame_steps = d_frame_steps;
        a.frames_out = frames_out;
        a.status_out = status_out;
        const size_t tmp_need = (size_t)count * n_frames * 4 * (size_t)(threads * cpt) * (elem2 / 2);
        a.frames_tmp = (tmp_base && n_frames > 0 && a.slot_plan && tmp_need <= tmp_avail && count <= 65535)
                           ? tmp_base : nullptr;
        if (a.frames_tmp && gs.concurrent()) tmp_cursor += align_up(tmp_need, 256);
        fn<<<count, threads, smem, stream>>>(a);
        CAB_CUDA(cudaGetLastError());
        if (a.frames_tmp) {
            const dim3 ug((unsigned)n_frames, (unsigned)count);
            if (precision == CAB200_FP64_STRICT)
                unpermute_kernel<double><<<ug, 256, 0, stream>>>((const double *)a.frames_tmp, (double *)frames_out, a.pouch_list,
                                                                 d_cell_off, a.slot_plan, n, 1, threads * cpt, n_frames);
            else
                unpermute_kernel<float><<<ug, 256, 0, stream>>>((const float *)a.frames_tmp, (float *)frames_out, a.pouch_list,
                                                                d_cell_off, a.slot_plan, n, 1, threads * cpt, n_frames);
            CAB_CUDA(cudaGetLastError());
        }
    }
    return CAB200_OK;
}
