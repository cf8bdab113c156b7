// prompts_host.cu -- host side of the point prompts for ALL frames of one pouch
// in one call, outside the interpreter.
//
// Replaces, per pouch of the batch path, the frame loop around the reference's
//   generate_point_prompts_from_mask   utils/sam2_data_generator.py:43-129
//   generate_negative_prompts          utils/sam2_data_generator.py:132-180
//   save_prompts                       utils/sam2_data_generator.py:210-240   (json.dump(indent=2) text)
// which calcium_b200/prompts.py (frame_prompts) restates frame by frame in numpy.  The frames of a pouch share one
// legacy numpy generator and must be drawn in order, and the Python version holds the interpreter lock for ~0.4 ms
// per frame (look-up tables, percentile, candidate selection, 1 400 formatted entries): with 13 320 frames per
// 148-pouch batch that lock, not the arithmetic, bounded the batch call.  Here one call per pouch does it all:
//   * the instance look-up table of the frame (core/pouch.py:540-565 + the reference's id mismatch, like
//     csrc/frame_lut.cuh);
//   * the two frame-dependent distance maps on the device (K4, cab200_edt_classes) on the caller's stream, copied
//     to pinned host memory;
//   * the draws, with numpy's own algorithms on the pouch's own Mersenne-Twister state: randint (masked rejection on
//     32-bit outputs, no draw for a range of one), choice(n, k, replace=False) = a Fisher-Yates shuffle (legacy_rng.cu),
//     np.percentile(..., 80) (method "linear": (m - 1) * 0.8, numpy's two-sided lerp);
//   * the prompt file's text, byte for byte what json.dump(indent=2) writes, written straight to its path.
// Same values, same generator state afterwards, same bytes as frame_prompts (tests/test_prompts.py).
#include <fcntl.h>
#include <math.h>
#include <unistd.h>

#include <algorithm>
#include <charconv>
#include <string>
#include <vector>

#include "common.cuh"

namespace cab200 {
// legacy_rng.cu
void legacy_choice_first_k(uint32_t *mt_key, int32_t *mt_pos, int64_t n, int32_t k, int64_t *out);
uint32_t legacy_next32(uint32_t *mt_key, int32_t *mt_pos);
}  // namespace cab200

using namespace cab200;

namespace {

// numpy's legacy bounded draw: value in [0, rng] by masked rejection on 32-bit outputs; no draw when rng == 0
inline uint32_t bounded32(uint32_t *key, int32_t *pos, uint32_t rng)
{
    if (rng == 0) return 0;
    uint32_t mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    uint32_t v;
    while ((v = (legacy_next32(key, pos) & mask)) > rng) {}
    return v;
}

// np.percentile(v, 80) for m >= 1 values given as squared integers (v = sqrt(d2), monotone): numpy partitions the
// array and interpolates between the two order statistics around (m - 1) * 0.8 with its two-sided lerp
double percentile80_of_sqrt(std::vector<int32_t> &d2)
{
    const int64_t m = (int64_t)d2.size();
    const double q = 80.0 / 100.0;
    const double virt = (double)(m - 1) * q;
    int64_t lo = (int64_t)floor(virt), hi = lo + 1;
    if (virt >= (double)(m - 1)) lo = hi = -1;
    double gamma = virt - (double)lo;
    if (lo < 0) lo = hi = m - 1;
    std::nth_element(d2.begin(), d2.begin() + lo, d2.end());
    const int32_t a2 = d2[(size_t)lo];
    int32_t b2 = a2;
    if (hi != lo) b2 = *std::min_element(d2.begin() + lo + 1, d2.end());
    const double a = sqrt((double)a2), b = sqrt((double)b2);
    const double diff = b - a;
    double r = a + diff * gamma;
    if (gamma >= 0.5) r = b - diff * (1.0 - gamma);
    return r;
}

inline void put_int(std::string &s, long long v)
{
    char b[24];
    auto r = std::to_chars(b, b + sizeof(b), v);
    s.append(b, r.ptr);
}

// float.__repr__ for the finite positive values met here (square roots of pixel distances): the shortest
// round-trip digits, with ".0" appended to integers
inline void put_float(std::string &s, double v)
{
    char b[40];
    auto r = std::to_chars(b, b + sizeof(b), v);
    bool plain = true;
    for (char *p = b; p < r.ptr; ++p) plain = plain && (*p >= '0' && *p <= '9');
    s.append(b, r.ptr);
    if (plain) s += ".0";
}

inline void put_point(std::string &s, long long x, long long y)
{
    s += "[\n        "; put_int(s, x); s += ",\n        "; put_int(s, y); s += "\n      ]";
}

bool write_file(const char *path, const std::string &text)
{
    const int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return false;
    const char *p = text.data();
    size_t left = text.size();
    while (left > 0) {
        const ssize_t w = write(fd, p, left);
        if (w < 0) { if (errno == EINTR) continue; close(fd); return false; }
        p += w; left -= (size_t)w;
    }
    close(fd);
    return true;
}

}  // namespace

extern "C" int cab200_np_percentile80_sqrt(const int32_t *d2, int64_t m, double *out)
{
    if (!d2 || !out || m < 1) { set_error("cab200_np_percentile80_sqrt: bad argument"); return CAB200_EINVAL; }
    std::vector<int32_t> v(d2, d2 + m);
    *out = percentile80_of_sqrt(v);
    return CAB200_OK;
}

extern "C" int cab200_legacy_randint(uint32_t *mt_key, int32_t *mt_pos, const int64_t *high, int64_t n, int64_t *out)
{
    if (!mt_key || !mt_pos || (n > 0 && (!high || !out)) || n < 0 || *mt_pos < 0 || *mt_pos > 624) {
        set_error("cab200_legacy_randint: bad argument");
        return CAB200_EINVAL;
    }
    for (int64_t i = 0; i < n; ++i) {
        if (high[i] < 1 || high[i] > 0xffffffffLL) { set_error("cab200_legacy_randint: high outside [1, 2^32]"); return CAB200_EINVAL; }
        out[i] = (int64_t)bounded32(mt_key, mt_pos, (uint32_t)(high[i] - 1));
    }
    return CAB200_OK;
}

extern "C" int cab200_prompts_pouch(const cab200_prompt_tables *t, int n_frames, const double *ca, double threshold,
                                    int quirk, int num_negative, int min_distance, int min_cell_area, uint32_t *mt_key,
                                    int32_t *mt_pos, const char *paths, const int64_t *path_off, void *dev_maps,
                                    void *dev_scratch, size_t scratch_bytes, void *dev_luts, void *host_maps,
                                    void *host_luts, void *stream_, int32_t *n_entries_out, char *text_out,
                                    int64_t text_cap, int64_t *text_off)
{
    if (!t || n_frames < 0 || !ca || !mt_key || !mt_pos || !dev_maps || !dev_scratch || !dev_luts || !host_maps ||
        !host_luts || !n_entries_out || *mt_pos < 0 || *mt_pos > 624 || num_negative < 0 || min_distance < 1 ||
        (text_out && !text_off)) {
        set_error("cab200_prompts_pouch: bad argument");
        return CAB200_EINVAL;
    }
    const int n = t->n_cells, H = t->H, W = t->W, L = n + 1;
    const int64_t npix = (int64_t)H * W;
    if (n < 1 || n > 65534 || H < 1 || W < 1 || !t->labels || !t->label_present || !t->area || !t->cand_off ||
        !t->cand_text || !t->cand_text_off || !t->label_map_dev || (t->empty_count > 0 && (!t->empty_flat || !t->empty_d2))) {
        set_error("cab200_prompts_pouch: incomplete tables");
        return CAB200_EINVAL;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    const int32_t min_d2 = min_distance * min_distance;
    uint16_t *hl = (uint16_t *)host_luts;          // [2][L] class tables (pinned)
    const int32_t *hm = (const int32_t *)host_maps;  // [2][H*W] distance maps (pinned)
    std::vector<uint16_t> lut(L);
    std::vector<int32_t> rank(n);
    std::vector<uint8_t> act(n);
    std::vector<std::pair<uint32_t, int32_t>> cells;       // (instance id, cell) of the frame's single-cell instances
    std::vector<int32_t> bgd2;
    std::vector<int64_t> bgpx;
    std::string text;
    int64_t text_used = 0;
    if (text_off) text_off[0] = 0;

    for (int f = 0; f < n_frames; ++f) {
        n_entries_out[f] = -1;                      // -1: no active cell, nothing is written (sam2_data_generator.py:318-319)
        if (text_off) text_off[f + 1] = text_used;
        const double *caf = ca + (size_t)f * n;
        // ---- instance id per label value: get_active_cells -> get_cell_masks(active_only) ->
        // create_instance_mask_from_cells, with the reference's comparison of 0-based ids with the 1-based mask
        int r = 0;
        bool any_active = false;
        for (int c = 0; c < n; ++c) {
            act[c] = caf[c] > threshold;
            any_active = any_active || act[c];
            rank[c] = act[c] ? ++r : 0;
        }
        if (!any_active) continue;
        if (quirk) {
            lut[0] = act[0] ? (uint16_t)rank[0] : 0;
            for (int v = 1; v <= n; ++v) {
                const int m = act[v - 1] ? v : 0;
                lut[v] = (m < n && act[m]) ? (uint16_t)rank[m] : 0;
            }
        } else {
            lut[0] = 0;
            for (int v = 1; v <= n; ++v) lut[v] = (uint16_t)rank[v - 1];
        }
        const uint32_t background_id = lut[0];           // id given to every zero pixel (cell 0 active), else 0
        bool any_labelled = false;
        for (int v = 0; v < L; ++v) any_labelled = any_labelled || (lut[v] && t->label_present[v]);
        // ---- the frame-dependent distance maps (K4): class (id == background id) and class (id > 0) ----------
        const int32_t *d2_bg = nullptr, *d2_neg = nullptr;
        if (background_id || (num_negative > 0 && any_labelled)) {
            const int k = background_id ? 2 : 1;
            if (background_id) {
                for (int v = 0; v < L; ++v) { hl[v] = lut[v] == background_id; hl[L + v] = lut[v] > 0; }
            } else {
                for (int v = 0; v < L; ++v) hl[v] = lut[v] > 0;
            }
            CAB_CUDA(cudaMemcpyAsync(dev_luts, hl, sizeof(uint16_t) * (size_t)k * L, cudaMemcpyHostToDevice, stream));
            if (int rc = cab200_edt_classes(t->label_map_dev, (const uint16_t *)dev_luts, L, k, H, W, (int32_t *)dev_maps,
                                            dev_scratch, scratch_bytes, stream_))
                return rc;
            CAB_CUDA(cudaMemcpyAsync(host_maps, dev_maps, sizeof(int32_t) * (size_t)k * npix, cudaMemcpyDeviceToHost, stream));
            CAB_CUDA(cudaStreamSynchronize(stream));
            if (background_id) { d2_bg = hm; d2_neg = hm + npix; } else { d2_neg = hm; }
        }
        // ---- positive prompts ----------------------------------------------------------------------------------
        cells.clear();
        for (int c = 0; c < n; ++c) {
            const uint32_t id = lut[c + 1];
            if (id && id != background_id && t->area[c] >= min_cell_area) cells.push_back({id, c});
        }
        std::stable_sort(cells.begin(), cells.end(), [](const std::pair<uint32_t, int32_t> &x, const std::pair<uint32_t, int32_t> &y) { return x.first < y.first; });
        text.clear();
        text += "{\n  \"positive_prompts\": ";
        const size_t pos_start = text.size();
        int n_entries = 0;
        auto open_entry = [&]() { text += n_entries ? ",\n" : "{\n"; ++n_entries; };
        size_t first_cell = 0;
        for (; first_cell < cells.size() && cells[first_cell].first < background_id; ++first_cell) {}   // (ids below the background's: none)
        auto draw_cells = [&](size_t lo, size_t hi) -> int {
            for (size_t i = lo; i < hi; ++i) {
                const int c = cells[i].second;
                const int64_t off = t->cand_off[c], cnt = t->cand_off[c + 1] - off;
                if (cnt < 1) { set_error("cab200_prompts_pouch: cell %d has no candidate pixel", c); return CAB200_EINVAL; }
                const int64_t pick = off + (int64_t)bounded32(mt_key, mt_pos, (uint32_t)(cnt - 1));
                open_entry();
                text += "    \""; put_int(text, cells[i].first); text += "\": ";
                text.append(t->cand_text + t->cand_text_off[pick], (size_t)(t->cand_text_off[pick + 1] - t->cand_text_off[pick]));
            }
            return CAB200_OK;
        };
        if (int rc = draw_cells(0, background_id ? first_cell : cells.size())) return rc;
        if (background_id) {
            // the "background" instance the id mismatch creates: every pixel whose label maps to that id.  The
            // reference's distance transform of (instance == id) is 0 outside the instance, so percentile and
            // candidates only ever see the instance's own pixels (row-major, the order np.where gives)
            bgd2.clear(); bgpx.clear();
            for (int64_t p = 0; p < npix; ++p)
                if (lut[t->labels[p]] == background_id) { bgpx.push_back(p); bgd2.push_back(d2_bg[p]); }
            const int64_t area = (int64_t)bgpx.size();
            if (area >= min_cell_area) {
                std::vector<int32_t> positive;
                positive.reserve(bgd2.size());
                for (int32_t v : bgd2) if (v > 0) positive.push_back(v);
                if (!positive.empty()) {                              // dist_transform.max() != 0
                    const double thr = percentile80_of_sqrt(positive);
                    int64_t n_cand = 0;
                    for (int32_t v : bgd2) n_cand += sqrt((double)v) >= thr;     // thr > 0
                    const int64_t idx = (int64_t)bounded32(mt_key, mt_pos, (uint32_t)(n_cand - 1));
                    int64_t seen = 0, px = -1;
                    int32_t v2 = 0;
                    for (size_t i = 0; i < bgd2.size(); ++i)
                        if (sqrt((double)bgd2[i]) >= thr) { if (seen == idx) { px = bgpx[i]; v2 = bgd2[i]; break; } ++seen; }
                    open_entry();
                    text += "    \""; put_int(text, background_id); text += "\": {\n      \"point\": ";
                    put_point(text, px % W, px / W);
                    text += ",\n      \"label\": 1,\n      \"area\": "; put_int(text, area);
                    text += ",\n      \"confidence\": "; put_float(text, sqrt((double)v2)); text += "\n    }";
                }
            }
            if (int rc = draw_cells(first_cell, cells.size())) return rc;
        }
        if (n_entries) text += "\n  }"; else text.insert(pos_start, "{}");
        // ---- negative prompts ----------------------------------------------------------------------------------
        text += ",\n  \"negative_prompts\": ";
        int n_neg = 0;
        int64_t negpx[64];
        double negconf[64];
        const int kneg = std::min(num_negative, 64);
        if (kneg > 0) {
            if (!any_labelled) {
                // no labelled pixel at all: scipy's distance_transform_edt of an array without a zero returns the
                // distance to (-1, 0); the reference samples its negatives from that map.  Static per output size.
                const int64_t cnt = t->empty_count;
                if (cnt > 0) {
                    int64_t pick[64];
                    n_neg = (int)std::min<int64_t>(kneg, cnt);
                    legacy_choice_first_k(mt_key, mt_pos, cnt, n_neg, pick);
                    for (int q = 0; q < n_neg; ++q) { negpx[q] = t->empty_flat[pick[q]]; negconf[q] = sqrt((double)t->empty_d2[negpx[q]]); }
                }
            } else {
                int64_t cnt = 0;
                for (int64_t p = 0; p < npix; ++p) cnt += (!lut[t->labels[p]] && d2_neg[p] >= min_d2) ? 1 : 0;
                if (cnt > 0) {
                    int64_t pick[64];
                    n_neg = (int)std::min<int64_t>(kneg, cnt);
                    legacy_choice_first_k(mt_key, mt_pos, cnt, n_neg, pick);
                    int order[64];
                    for (int q = 0; q < n_neg; ++q) order[q] = q;
                    std::sort(order, order + n_neg, [&](int x, int y) { return pick[x] < pick[y]; });
                    int64_t seen = 0;
                    int next = 0;
                    for (int64_t p = 0; p < npix && next < n_neg; ++p)
                        if (!lut[t->labels[p]] && d2_neg[p] >= min_d2) {
                            while (next < n_neg && pick[order[next]] == seen) {
                                negpx[order[next]] = p; negconf[order[next]] = sqrt((double)d2_neg[p]); ++next;
                            }
                            ++seen;
                        }
                }
            }
        }
        if (n_neg) {
            text += "[\n";
            for (int q = 0; q < n_neg; ++q) {
                if (q) text += ",\n";
                text += "    {\n      \"point\": "; put_point(text, negpx[q] % W, negpx[q] / W);
                text += ",\n      \"label\": 0,\n      \"confidence\": "; put_float(text, negconf[q]); text += "\n    }";
            }
            text += "\n  ]";
        } else {
            text += "[]";
        }
        text += ",\n  \"num_cells\": "; put_int(text, n_entries); text += ",\n  \"version\": \"1.0\"\n}";
        n_entries_out[f] = n_entries;
        if (paths && path_off && path_off[f + 1] > path_off[f] && paths[path_off[f]]) {
            if (!write_file(paths + path_off[f], text)) {
                set_error("cab200_prompts_pouch: cannot write %s: %s", paths + path_off[f], strerror(errno));
                return CAB200_EINVAL;
            }
        }
        if (text_out) {
            if (text_used + (int64_t)text.size() > text_cap) { set_error("cab200_prompts_pouch: text buffer too small"); return CAB200_ESIZE; }
            memcpy(text_out + text_used, text.data(), text.size());
            text_used += (int64_t)text.size();
            text_off[f + 1] = text_used;
        }
    }
    return CAB200_OK;
}
