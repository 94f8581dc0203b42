// individual_image_normalization = moments ("next" row N3b): the blob is rotated by its second-moment orientation into the
// 80x80 canvas.  Replaces, for grey blobs,
//   pv::Blob::calculate_moments              C/processing/PVBlob.cpp:111-214 (incl. its four-package float sums beyond 1000 runs)
//   fast_atan2                               C/misc/math.h:34-59
//   constraints::diff_image (moments branch) T/tracking/FilterCache.cpp:329-341
//   image::normalize_image                   T/tracking/FilterCache.cpp:21-115 (gui::Transform in double, C/gui/Transform.cpp)
//   cv::warpAffine(INTER_LINEAR, BORDER_CONSTANT) for 8-bit single-channel images (OpenCV 4.x fixed point: coordinates
//   with 10 fractional bits of which 5 are kept, 15-bit tap weights (32-fx)(32-fy)*32, +2^14 >> 15)
// This file is compiled with -fmad=false: the float accumulations, the polynomial and the double matrix algebra must round
// exactly like the reference's scalar code (no fused multiply-adds).
#include "common.h"

#include <math.h>

namespace tb {

__device__ __forceinline__ float fast_atan_f(float z) { const float n1 = 0.97239411f, n2 = -0.19194795f; return (n1 + n2 * z * z) * z; }
__device__ __forceinline__ float fast_atan2_f(float y, float x)
{
    if (x == 0.0f) return copysignf((float)1.57079632679489661923, y);
    const float abs_y = fabsf(y);
    float r, angle;
    if (abs_y < fabsf(x)) { r = abs_y / fabsf(x); angle = fast_atan_f(r); }
    else { r = fabsf(x) / abs_y; angle = (float)(1.57079632679489661923 - (double)fast_atan_f(r)); }
    if (x < 0.0f) angle = (float)(3.14159265358979323846 - (double)angle);
    if (y < 0.0f) angle = -angle;
    return angle;
}

// One thread per crop: moments in the reference's pixel order, orientation, forward matrix, and the INVERTED matrix
// cv::warpAffine works with (coef[0..5]).
__global__ void blob_moments_kernel(const tb_blob_rec *__restrict__ recs, const uint32_t *__restrict__ totals, const uint32_t *__restrict__ crop_blob,
                                    const tb_line *__restrict__ lines, int out_w, int out_h, double *__restrict__ coef)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= totals[3]) return;
    const tb_blob_rec r = recs[crop_blob[q]];
    const tb_line *L = lines + r.line_off;
    // the reference sums the runs in `packages` consecutive packages with their own float accumulators and merges them in order: one package up to 1000
    // runs, min(runs, 4) beyond (PVBlob.cpp:118-122,190-204; distribute_indexes, C/misc/ThreadPool.h:126-195: runs / packages each, the last takes the rest)
    const uint32_t packages = r.n_lines > 1000u ? 4u : 1u;
    const uint32_t per = r.n_lines / packages;
    float m00 = 0.f, m01 = 0.f, m10 = 0.f;
    for (uint32_t p = 0; p < packages; ++p) {
        const uint32_t b = p * per, e = p + 1 == packages ? r.n_lines : (p + 1) * per;
        float l00 = 0.f, l01 = 0.f, l10 = 0.f;
        for (uint32_t i = b; i < e; ++i) {
            const tb_line l = L[i];
            const float my = (float)(unsigned)l.y;
            for (int x = l.x0; x <= (int)l.x1; ++x) { l00 += 1.f; l01 += my; l10 += (float)x; }
        }
        m00 += l00; m01 += l01; m10 += l10;
    }
    const float cx = m10 / m00, cy = m01 / m00;
    float mu00 = 0.f, mu02 = 0.f, mu11 = 0.f, mu20 = 0.f;
    for (uint32_t p = 0; p < packages; ++p) {
        const uint32_t b = p * per, e = p + 1 == packages ? r.n_lines : (p + 1) * per;
        float l00 = 0.f, l02 = 0.f, l11 = 0.f, l20 = 0.f;
        for (uint32_t i = b; i < e; ++i) {
            const tb_line l = L[i];
            const int vy = (int)((float)l.y - cy);
            const int vy2 = vy * vy;
            int vx = (int)((float)l.x0 - cx);
            for (int x = l.x0; x <= (int)l.x1; ++x, ++vx) {
                l00 += 1.f; l02 += (float)vy2; l11 += (float)vx * (float)vy; l20 += (float)(vx * vx);
            }
        }
        mu00 += l00; mu02 += l02; mu11 += l11; mu20 += l20;
    }
    const float inv = 1.0f / mu00;
    const float orientation = (float)(0.5 * (double)fast_atan2_f(2 * (mu11 * inv), mu20 * inv - mu02 * inv));
    // FilterCache.cpp:333-337 + normalize_image:47-62
    const float deg = (-orientation + 3.14159274f * 0.25f) * (1.0f / 3.14159274f * 180.0f);
    const double rad = (double)deg * 3.141592654 / 180.0;
    const double c = cos(rad), s = sin(rad);
    const int bw = (int)r.x1 - (int)r.x0 + 1, bh = (int)r.y1 - (int)r.y0 + 1;
    const double tx = (double)(-((float)bw * 0.5f)), ty = (double)(-((float)bh * 0.5f));
    const double r02 = c * tx + (-s) * ty + 0.0, r12 = s * tx + c * ty + 0.0;
    double M[6];
    const double ox = (double)((float)out_w * 0.5f), oy = (double)((float)out_h * 0.5f);
    M[0] = c; M[1] = -s; M[2] = r02 + ox;
    M[3] = s; M[4] = c;  M[5] = r12 + oy;
    // cv::warpAffine: invert (imgwarp.cpp)
    double D = M[0] * M[4] - M[1] * M[3];
    D = D != 0 ? 1. / D : 0;
    const double A11 = M[4] * D, A22 = M[0] * D;
    M[0] = A11; M[1] *= -D; M[3] *= -D; M[4] = A22;
    const double b1 = -M[0] * M[2] - M[1] * M[5], b2 = -M[3] * M[2] - M[4] * M[5];
    M[2] = b1; M[5] = b2;
    for (int k = 0; k < 6; ++k) coef[(size_t)q * 6 + k] = M[k];
}

// One CTA per crop: every output pixel takes four taps of the blob's bounding-box image (difference image under the
// mask, zero elsewhere), which is never materialised: a tap finds its line through a per-row index in shared memory.
constexpr int CW_NT = 256, CW_ROWS = 1024;

__global__ void __launch_bounds__(CW_NT)
crop_warp_kernel(const tb_blob_rec *__restrict__ recs, const uint32_t *__restrict__ totals, const uint32_t *__restrict__ crop_blob,
                 const tb_line *__restrict__ lines, const uint32_t *__restrict__ line_px, const uint8_t *__restrict__ pixels,
                 const uint8_t *__restrict__ bg, int W, int crop_method, int out_w, int out_h,
                 const double *__restrict__ coef, uint8_t *__restrict__ crops)
{
    __shared__ uint16_t s_first[CW_ROWS], s_cnt[CW_ROWS];
    const uint32_t q = blockIdx.x;
    if (q >= totals[3]) return;
    const tb_blob_rec r = recs[crop_blob[q]];
    const tb_line *L = lines + r.line_off;
    const uint32_t *LP = line_px + r.line_off;
    const int bw = (int)r.x1 - (int)r.x0 + 1, bh = (int)r.y1 - (int)r.y0 + 1;
    const bool indexed = bh <= CW_ROWS && r.n_lines < 65536u;
    if (indexed) {
        for (int i = threadIdx.x; i < bh; i += CW_NT) { s_first[i] = 0; s_cnt[i] = 0; }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < r.n_lines; i += CW_NT) {
            const int yy = (int)L[i].y - (int)r.y0;
            if (i == 0 || L[i - 1].y != L[i].y) s_first[yy] = (uint16_t)i;
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < r.n_lines; i += CW_NT) {
            const int yy = (int)L[i].y - (int)r.y0;
            if (i + 1 == r.n_lines || L[i + 1].y != L[i].y) s_cnt[yy] = (uint16_t)(i + 1 - s_first[yy]);
        }
        __syncthreads();
    }
    double M[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) M[k] = coef[(size_t)q * 6 + k];
    uint8_t *out = crops + (size_t)q * out_w * out_h;
    auto tap = [&](int yy, int xx) -> int {
        if (yy < 0 || yy >= bh || xx < 0 || xx >= bw) return 0;
        const int ax = xx + (int)r.x0, ay = yy + (int)r.y0;
        uint32_t j0, j1;
        if (indexed) { j0 = s_first[yy]; j1 = j0 + s_cnt[yy]; }
        else {                                               // first line of row ay by binary search
            uint32_t lo = 0, hi = r.n_lines;
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if ((int)L[mid].y < ay) lo = mid + 1; else hi = mid; }
            j0 = lo; j1 = r.n_lines;
        }
        for (uint32_t j = j0; j < j1; ++j) {
            const tb_line l = L[j];
            if ((int)l.y != ay || (int)l.x0 > ax) break;
            if (ax <= (int)l.x1) {
                int v = pixels[LP[j] + (uint32_t)(ax - (int)l.x0)];
                if (crop_method) {
                    const int b = bg[(size_t)ay * W + ax];
                    v = crop_method == 1 ? abs(b - v) : max(0, b - v);
                }
                return v;
            }
        }
        return 0;
    };
    for (int i = threadIdx.x; i < out_w * out_h; i += CW_NT) {
        const int y = i / out_w, x = i % out_w;
        const int X0 = __double2int_rn((M[1] * y + M[2]) * 1024) + 16, Y0 = __double2int_rn((M[4] * y + M[5]) * 1024) + 16;
        const int X = (X0 + __double2int_rn(M[0] * x * 1024)) >> 5, Y = (Y0 + __double2int_rn(M[3] * x * 1024)) >> 5;
        const int sx = X >> 5, sy = Y >> 5, fx = X & 31, fy = Y & 31;
        int acc = 0;
        if (sx >= -1 && sx < bw && sy >= -1 && sy < bh) {
            acc = tap(sy, sx) * ((32 - fx) * (32 - fy) * 32) + tap(sy, sx + 1) * (fx * (32 - fy) * 32) +
                  tap(sy + 1, sx) * ((32 - fx) * fy * 32) + tap(sy + 1, sx + 1) * (fx * fy * 32);
        }
        out[i] = (uint8_t)((acc + (1 << 14)) >> 15);
    }
}

// individual_image_scale != 1 (T/tracking/FilterCache.cpp:178-180): the masked blob image goes through
// resize_image = cv::resize(INTER_NEAREST) (C/misc/detail.h:465-469: dsize = cvRound(n * f), source = min(floor(d / f), n - 1))
// before the centre pad / centre crop (:184-228).  One CTA per crop gathers the output pixels; a pixel finds its value
// through the blob's rows like crop_warp_kernel does.
__global__ void __launch_bounds__(CW_NT)
crop_scale_kernel(const tb_blob_rec *__restrict__ recs, const uint32_t *__restrict__ totals, const uint32_t *__restrict__ crop_blob,
                  const tb_line *__restrict__ lines, const uint32_t *__restrict__ line_px, const uint8_t *__restrict__ pixels,
                  const uint8_t *__restrict__ bg, int W, int crop_method, int out_w, int out_h, float scale, uint8_t *__restrict__ crops)
{
    __shared__ uint16_t s_first[CW_ROWS];
    const uint32_t q = blockIdx.x;
    if (q >= totals[3]) return;
    const tb_blob_rec r = recs[crop_blob[q]];
    const tb_line *L = lines + r.line_off;
    const uint32_t *LP = line_px + r.line_off;
    const int bw = (int)r.x1 - (int)r.x0 + 1, bh = (int)r.y1 - (int)r.y0 + 1;
    const bool indexed = bh <= CW_ROWS && r.n_lines < 65536u;
    if (indexed) {
        for (uint32_t i = threadIdx.x; i < r.n_lines; i += CW_NT)
            if (i == 0 || L[i - 1].y != L[i].y) s_first[(int)L[i].y - (int)r.y0] = (uint16_t)i;      // a connected blob has no empty row
        __syncthreads();
    }
    const double f = (double)scale, inv = 1.0 / f;
    const int dw = __double2int_rn((double)bw * f), dh = __double2int_rn((double)bh * f);            // cvRound
    int offx, offy;                                  // output = resized + offset (pad: >= 0, crop: < 0)
    if (dw < out_w) { const int d = out_w - dw; offx = d - d / 2; } else { const int d = dw - out_w; offx = -(d - d / 2); }
    if (dh < out_h) { const int d = out_h - dh; offy = d - d / 2; } else { const int d = dh - out_h; offy = -(d - d / 2); }
    uint8_t *out = crops + (size_t)q * out_w * out_h;
    for (int i = threadIdx.x; i < out_w * out_h; i += CW_NT) {
        const int rx = i % out_w - offx, ry = i / out_w - offy;
        int v = 0;
        if (rx >= 0 && rx < dw && ry >= 0 && ry < dh) {
            const int sx = min((int)floor((double)rx * inv), bw - 1), sy = min((int)floor((double)ry * inv), bh - 1);
            const int ax = sx + (int)r.x0, ay = sy + (int)r.y0;
            uint32_t j0;
            if (indexed) j0 = s_first[sy];
            else {
                uint32_t lo = 0, hi = r.n_lines;
                while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if ((int)L[mid].y < ay) lo = mid + 1; else hi = mid; }
                j0 = lo;
            }
            for (uint32_t j = j0; j < r.n_lines; ++j) {
                const tb_line l = L[j];
                if ((int)l.y != ay || (int)l.x0 > ax) break;
                if (ax <= (int)l.x1) {
                    v = pixels[LP[j] + (uint32_t)(ax - (int)l.x0)];
                    if (crop_method) {
                        const int b = bg[(size_t)ay * W + ax];
                        v = crop_method == 1 ? abs(b - v) : max(0, b - v);
                    }
                    break;
                }
            }
        }
        out[i] = (uint8_t)v;
    }
}

int launch_crop_scaled(const tb_blob_rec *recs, const uint32_t *totals, const uint32_t *crop_blob, const tb_line *lines,
                       const uint32_t *line_px, const uint8_t *pixels, const uint8_t *bg, int W, int crop_method,
                       int out_w, int out_h, float scale, uint8_t *crops, int max_crops_total, cudaStream_t s)
{
    if (max_crops_total <= 0) return TB_OK;
    crop_scale_kernel<<<max_crops_total, CW_NT, 0, s>>>(recs, totals, crop_blob, lines, line_px, pixels, bg, W, crop_method, out_w, out_h, scale, crops);
    TB_CUDA(cudaGetLastError());
    return TB_OK;
}

int launch_crop_warp(const tb_blob_rec *recs, const uint32_t *totals, const uint32_t *crop_blob, const tb_line *lines,
                     const uint32_t *line_px, const uint8_t *pixels, const uint8_t *bg, int W, int crop_method,
                     int out_w, int out_h, const double *coef, uint8_t *crops, int max_crops_total, cudaStream_t s)
{
    if (max_crops_total <= 0) return TB_OK;
    crop_warp_kernel<<<max_crops_total, CW_NT, 0, s>>>(recs, totals, crop_blob, lines, line_px, pixels, bg, W, crop_method, out_w, out_h, coef, crops);
    TB_CUDA(cudaGetLastError());
    return TB_OK;
}

int launch_crop_moments(const tb_blob_rec *recs, const uint32_t *totals, const uint32_t *crop_blob, const tb_line *lines,
                        const uint32_t *line_px, const uint8_t *pixels, const uint8_t *bg, int W, int crop_method,
                        int out_w, int out_h, uint8_t *crops, double *coef, int max_crops_total, cudaStream_t s)
{
    if (max_crops_total <= 0) return TB_OK;
    blob_moments_kernel<<<(max_crops_total + 127) / 128, 128, 0, s>>>(recs, totals, crop_blob, lines, out_w, out_h, coef);
    crop_warp_kernel<<<max_crops_total, CW_NT, 0, s>>>(recs, totals, crop_blob, lines, line_px, pixels, bg, W, crop_method, out_w, out_h, coef, crops);
    TB_CUDA(cudaGetLastError());
    return TB_OK;
}

}  // namespace tb
