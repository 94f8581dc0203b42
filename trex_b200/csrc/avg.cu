// Background image generation on the GPU ("next" row N2b): cmn::AveragingAccumulator
// (C/video/AveragingAccumulator.cpp:23-196), the step before the segmentation path
// (VideoSource::generate_average, C/video/VideoSource.cpp:940-1030 -> BackgroundSubtraction::set_background).
//   mean: float sum, finalize = round-half-even(sum / count) saturated to u8   (:63-73,159-161)
//   mode: per-pixel histogram, finalize = smallest value with the highest count (:75-131,163-190)
//   max / min: element-wise                                                    (:133-148)
// One thread per pixel; frames of an add() call are walked in registers (u8 sums are exact in fp32).
#include "common.h"

#include <vector>

namespace tb {

enum { AVG_MEAN = 0, AVG_MODE = 1, AVG_MAX = 2, AVG_MIN = 3 };

__global__ void avg_add_kernel(const uint8_t *__restrict__ frames, int n, size_t px, int method,
                               float *__restrict__ fsum, uint8_t *__restrict__ ext, uint16_t *__restrict__ hist)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < px; i += (size_t)gridDim.x * blockDim.x) {
        if (method == AVG_MEAN) {
            float s = fsum[i];
            for (int f = 0; f < n; ++f) s += (float)frames[(size_t)f * px + i];
            fsum[i] = s;
        } else if (method == AVG_MODE) {
            for (int f = 0; f < n; ++f) hist[(size_t)frames[(size_t)f * px + i] * px + i] += 1;   // [bin][pixel]: the thread owns its column
        } else {
            int v = ext[i];
            for (int f = 0; f < n; ++f) {
                const int q = frames[(size_t)f * px + i];
                v = method == AVG_MAX ? max(v, q) : min(v, q);
            }
            ext[i] = (uint8_t)v;
        }
    }
}

__global__ void avg_finalize_kernel(size_t px, int method, double count, const float *__restrict__ fsum,
                                    const uint8_t *__restrict__ ext, const uint16_t *__restrict__ hist, uint8_t *__restrict__ out)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < px; i += (size_t)gridDim.x * blockDim.x) {
        if (method == AVG_MEAN) {
            const float q = fsum[i] / (float)count;                     // cv::divide on CV_32F: true division
            out[i] = (uint8_t)min(max(__float2int_rn(q), 0), 255);      // convertTo: cvRound (half to even) + saturate
        } else if (method == AVG_MODE) {
            int best = 0; unsigned bc = hist[i];
            for (int b = 1; b < 256; ++b) { const unsigned c = hist[(size_t)b * px + i]; if (c > bc) { bc = c; best = b; } }   // first maximum
            out[i] = (uint8_t)best;
        } else out[i] = ext[i];
    }
}

}  // namespace tb

using namespace tb;

struct tb_avg {
    int device = 0, w = 0, h = 0, method = 0;
    size_t px = 0;
    double count = 0;
    float *fsum = nullptr; uint8_t *ext = nullptr; uint16_t *hist = nullptr;
    uint8_t *stage = nullptr, *out = nullptr; int stage_frames = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_last = nullptr; bool ev_valid = false;    // completion of the last accumulate kernel, whatever stream it ran on
    uint64_t launches = 0;
};

extern "C" int tb_avg_create(int device, int width, int height, int method, tb_avg **out)
{
    TB_REQUIRE(out && width > 0 && height > 0, TB_ERR_INVALID, "tb_avg_create: bad argument");
    TB_REQUIRE(method >= 0 && method <= 3, TB_ERR_INVALID, "tb_avg_create: method must be 0 mean, 1 mode, 2 max, 3 min");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("tb_avg_create: no CUDA device (there is no CPU fallback)"); return TB_ERR_CUDA; }
    TB_REQUIRE(device >= 0 && device < ndev, TB_ERR_INVALID, "tb_avg_create: bad device ordinal");
    TB_CUDA(cudaSetDevice(device));
    tb_avg *h = new tb_avg();
    h->device = device; h->w = width; h->h = height; h->method = method; h->px = (size_t)width * height;
    h->stage_frames = 16;
    int r = TB_OK;
    if (method == AVG_MEAN) { r = dev_alloc(&h->fsum, h->px); if (r == TB_OK) cudaMemset(h->fsum, 0, h->px * 4); }
    else if (method == AVG_MODE) { r = dev_alloc(&h->hist, h->px * 256); if (r == TB_OK) cudaMemset(h->hist, 0, h->px * 512); }
    else { r = dev_alloc(&h->ext, h->px); if (r == TB_OK) cudaMemset(h->ext, method == AVG_MIN ? 255 : 0, h->px); }
    if (r == TB_OK) r = dev_alloc(&h->stage, h->px * h->stage_frames);
    if (r == TB_OK) r = dev_alloc(&h->out, h->px);
    if (r == TB_OK && cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); r = TB_ERR_CUDA; }
    if (r == TB_OK && cudaEventCreateWithFlags(&h->ev_last, cudaEventDisableTiming) != cudaSuccess) { set_error("cudaEventCreate failed"); r = TB_ERR_CUDA; }
    if (r != TB_OK) { tb_avg_destroy(h); return r; }
    *out = h;
    return TB_OK;
}

extern "C" void tb_avg_destroy(tb_avg *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    cudaFree(h->fsum); cudaFree(h->ext); cudaFree(h->hist); cudaFree(h->stage); cudaFree(h->out);
    if (h->ev_last) cudaEventDestroy(h->ev_last);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" int tb_avg_add_device(tb_avg *h, const void *frames_dev, int n, void *stream)
{
    TB_REQUIRE(h && frames_dev && n > 0, TB_ERR_INVALID, "tb_avg_add_device: bad argument");
    TB_REQUIRE(h->method != AVG_MODE || h->count + n <= 65535, TB_ERR_CAPACITY, "tb_avg_add_device: mode histogram counters hold at most 65535 samples");
    TB_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
    // the accumulators are read-modify-written: every add (and the finalize) is ordered behind the previous add, also across streams
    if (h->ev_valid) TB_CUDA(cudaStreamWaitEvent(s, h->ev_last, 0));
    avg_add_kernel<<<148 * 8, 256, 0, s>>>((const uint8_t *)frames_dev, n, h->px, h->method, h->fsum, h->ext, h->hist);
    TB_CUDA(cudaGetLastError());
    TB_CUDA(cudaEventRecord(h->ev_last, s));
    h->ev_valid = true;
    h->count += n; h->launches += 1;
    return TB_OK;
}

extern "C" int tb_avg_add(tb_avg *h, const uint8_t *frames, int n)
{
    TB_REQUIRE(h && frames && n > 0, TB_ERR_INVALID, "tb_avg_add: bad argument");
    TB_CUDA(cudaSetDevice(h->device));
    for (int i = 0; i < n; i += h->stage_frames) {
        const int m = std::min(h->stage_frames, n - i);
        TB_CUDA(cudaMemcpyAsync(h->stage, frames + (size_t)i * h->px, (size_t)m * h->px, cudaMemcpyHostToDevice, h->stream));
        int r = tb_avg_add_device(h, h->stage, m, h->stream);
        if (r != TB_OK) return r;
        TB_CUDA(cudaStreamSynchronize(h->stream));
    }
    return TB_OK;
}

extern "C" int tb_avg_finalize(tb_avg *h, uint8_t *out_host)
{
    TB_REQUIRE(h && out_host, TB_ERR_INVALID, "tb_avg_finalize: null argument");
    TB_REQUIRE(h->count > 0, TB_ERR_STATE, "tb_avg_finalize: no samples added");
    TB_CUDA(cudaSetDevice(h->device));
    if (h->ev_valid) TB_CUDA(cudaStreamWaitEvent(h->stream, h->ev_last, 0));
    avg_finalize_kernel<<<148 * 8, 256, 0, h->stream>>>(h->px, h->method, h->count, h->fsum, h->ext, h->hist, h->out);
    TB_CUDA(cudaGetLastError());
    h->launches += 1;
    TB_CUDA(cudaMemcpyAsync(out_host, h->out, h->px, cudaMemcpyDeviceToHost, h->stream));
    TB_CUDA(cudaStreamSynchronize(h->stream));
    return TB_OK;
}
