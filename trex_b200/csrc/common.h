// Shared helpers for libtrexb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>

#include "../../include/trexb200.h"

namespace tb {

void set_error(const std::string &msg);
const char *get_error();

#define TB_CUDA(expr)                                                                        \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            tb::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));               \
            return TB_ERR_CUDA;                                                              \
        }                                                                                    \
    } while (0)

#define TB_REQUIRE(cond, code, msg)                                                          \
    do {                                                                                     \
        if (!(cond)) { tb::set_error(msg); return (code); }                                  \
    } while (0)

template <typename T>
static inline int dev_alloc(T **p, size_t n)
{
    cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
    if (e != cudaSuccess) { set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e)); return TB_ERR_CUDA; }
    return TB_OK;
}
template <typename T>
static inline int host_alloc(T **p, size_t n)
{
    cudaError_t e = cudaMallocHost((void **)p, n * sizeof(T));
    if (e != cudaSuccess) { set_error(std::string("cudaMallocHost: ") + cudaGetErrorString(e)); return TB_ERR_CUDA; }
    return TB_OK;
}

// cudaFuncSetAttribute is per device: a call site keeps one of these (static) and repeats its setup on every device it meets
struct DeviceOnce {
    std::atomic<unsigned long long> mask{0};
    static int dev() { int d = 0; cudaGetDevice(&d); return d & 63; }
    bool need() const { return !((mask.load(std::memory_order_acquire) >> dev()) & 1ull); }
    void done() { mask.fetch_or(1ull << dev(), std::memory_order_release); }
};

// Ring of CUDA-event brackets around the NK kernels of one launch sequence (measurement hook).
template <int NK>
struct EventRing {
    static constexpr int SLOTS = 32;
    cudaEvent_t ev[SLOTS][NK + 1] = {};
    cudaStream_t stream[SLOTS] = {};
    int used = 0;
    bool enabled = false, created = false;
    double acc[NK] = {};
    uint64_t n = 0;
    int enable(bool on)
    {
        if (on && !created) {
            for (auto &row : ev)
                for (auto &e : row)
                    if (cudaEventCreate(&e) != cudaSuccess) return TB_ERR_CUDA;
            created = true;
        }
        enabled = on;
        return TB_OK;
    }
    int flush()
    {
        for (int s = 0; s < used; ++s) {
            if (cudaEventSynchronize(ev[s][NK]) != cudaSuccess) return TB_ERR_CUDA;
            for (int k = 0; k < NK; ++k) {
                float ms = 0.f;
                if (cudaEventElapsedTime(&ms, ev[s][k], ev[s][k + 1]) != cudaSuccess) return TB_ERR_CUDA;
                acc[k] += ms;
            }
            ++n;
        }
        used = 0;
        return TB_OK;
    }
    // returns the slot to record into, or -1 when profiling is off
    int begin(cudaStream_t st)
    {
        if (!enabled) return -1;
        if (used == SLOTS) flush();
        stream[used] = st;
        return used++;
    }
    void mark(int slot, int k) { if (slot >= 0) cudaEventRecord(ev[slot][k], stream[slot]); }
    void destroy()
    {
        if (created) for (auto &row : ev) for (auto &e : row) cudaEventDestroy(e);
        created = false;
    }
};

// crop_norm.cu: `moments` normalisation of the crops of a batch (moments + matrix per crop, then the warp); 2 launches
int launch_crop_moments(const tb_blob_rec *recs, const uint32_t *totals, const uint32_t *crop_blob, const tb_line *lines,
                        const uint32_t *line_px, const uint8_t *pixels, const uint8_t *bg, int W, int crop_method,
                        int out_w, int out_h, uint8_t *crops, double *coef, int max_crops_total, cudaStream_t s);

// crop_norm.cu: crops with individual_image_scale != 1 (nearest-neighbour resize before the centre pad / crop); 1 launch
int launch_crop_scaled(const tb_blob_rec *recs, const uint32_t *totals, const uint32_t *crop_blob, const tb_line *lines,
                       const uint32_t *line_px, const uint8_t *pixels, const uint8_t *bg, int W, int crop_method,
                       int out_w, int out_h, float scale, uint8_t *crops, int max_crops_total, cudaStream_t s);

// box.cu: cv::boxFilter / cv::blur of n u8 planes (k x k mean, border 0 replicate / 1 reflect-101); hs = scratch of sub*W*H words;
// 2 launches per sub-batch of `sub` frames
int launch_box_mean(const uint8_t *src, uint8_t *dst, uint32_t *hs, int sub, int W, int H, int n, int k, int border, cudaStream_t s);

// crop_norm.cu: cv::warpAffine of every crop's blob image with the inverted map coef[6 * crop]; 1 launch
int launch_crop_warp(const tb_blob_rec *recs, const uint32_t *totals, const uint32_t *crop_blob, const tb_line *lines,
                     const uint32_t *line_px, const uint8_t *pixels, const uint8_t *bg, int W, int crop_method,
                     int out_w, int out_h, const double *coef, uint8_t *crops, int max_crops_total, cudaStream_t s);

// outline.cu: longest outline (pixel::find_outer_points) of the batch's blobs + Outline::resample; 1 memset + 3 launches.
// nb_dev: device word holding the number of blobs (nullptr: nb_max is the count); nb_max bounds the grids.
// map (optional): record q takes its outline from blob index[q] (0xFFFFFFFF: none) with points relative to origin[q]'s bounds; records with
// skip[q] != 0 are left alone; append: the point arenas continue behind totals[0..1] (posture of thresholded sub-blobs, round by round)
struct OutlineMap { const uint32_t *index; const tb_blob_rec *origin; const uint8_t *skip; int append; };
int launch_outlines(const tb_blob_rec *recs, const uint32_t *nb_dev, uint32_t nb_max, const tb_line *lines, const uint32_t *line_px, int opx,
                    uint8_t *visited, size_t visited_bytes, float rd, uint32_t *row_first, int4 *sel,
                    tb_outline_rec *orecs, uint32_t *totals, float *raw, float *res, uint32_t cap_pts, int sms, cudaStream_t s,
                    const OutlineMap *map = nullptr);

// posture.cu: Outline::calculate_midline (+ Midline::post_process / normalize with do_norm) for the batch's resampled outlines; 1 memset + 1 launch
int launch_midlines(const tb_outline_rec *orecs, const uint32_t *nb_dev, uint32_t nb_max, const float *res, uint32_t cap_pts,
                    const tb_posture_params *P, int do_norm, const float *move_dir, const float *fix_len,
                    float *pts_out, float *segs, tb_midline_rec *mrecs, tb_midline_norm *nrecs, float *norm_pts,
                    float *arena, unsigned long long arena_floats, unsigned long long *arena_used, uint32_t *status, int sms, cudaStream_t s,
                    const uint8_t *skip = nullptr);

// posture.cu: posture::calculate_posture's threshold loop, per round: parents of the re-thresholded sub-blobs + the biggest one per parent,
// and the per-parent state update after the round's midlines
struct PostureRound {
    const tb_blob_rec *parent_recs; const tb_frame_info *parent_infos; const tb_line *parent_lines; const uint32_t *n_parents;   // the blobs whose posture is wanted
    const tb_blob_rec *sub_recs; const tb_line *sub_lines; const uint32_t *n_subs;                                             // their sub-blobs at this round's threshold
    unsigned long long *best; uint32_t *index; uint32_t *sub_npx; uint8_t *state; tb_outline_rec *first_outline;               // per parent
    uint32_t *remaining;
};
int launch_posture_parents(const PostureRound &R, uint32_t max_parents, uint32_t max_subs, int sms, cudaStream_t s);
int launch_posture_round_end(const PostureRound &R, uint32_t max_parents, tb_outline_rec *orecs, tb_midline_rec *mrecs, tb_midline_norm *nrecs,
                             int do_norm, int last_round, int sms, cudaStream_t s);

// posture.cu: `posture` / `legacy` crops from the normalised midlines (map per crop, then crop_norm.cu's warp); 2 launches
int launch_posture_crops(const tb_blob_rec *recs, const uint32_t *totals, const uint32_t *crop_blob, const tb_line *lines,
                         const uint32_t *line_px, const uint8_t *pixels, const uint8_t *bg, int W, int crop_method,
                         int out_w, int out_h, const tb_midline_norm *nrecs, const float *median_len, float median_len_all,
                         float image_scale, int legacy, uint8_t *crops, double *coef, uint8_t *valid, int max_crops_total, cudaStream_t s);

#ifdef __CUDACC__
// Exclusive scan of one value per thread across the CTA; `total` = sum over all threads.
// ws: shared array of >= 33 uint32. All threads must call.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *ws, uint32_t &total)
{
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nw = (blockDim.x + 31u) >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (unsigned)d) inc += t;
    }
    __syncthreads();                      // protect ws from a previous use
    if (lane == 31) ws[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < nw ? ws[lane] : 0u, winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= (unsigned)d) winc += t;
        }
        if (lane < nw) ws[lane] = winc - w;
        if (lane == 31) ws[32] = winc;
    }
    __syncthreads();
    total = ws[32];
    return ws[warp] + inc - v;
}

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, uint32_t &total)
{
    const unsigned lane = threadIdx.x & 31u;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (unsigned)d) inc += t;
    }
    total = __shfl_sync(0xffffffffu, inc, 31);
    return inc - v;
}
#endif

}  // namespace tb
