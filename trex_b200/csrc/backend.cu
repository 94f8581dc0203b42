// Plug-in entry of libtrexb200.so: trex_b200_register() + the pure-C back-end table (include/trexb200.h).
//
// Mirrors the reference's own plug-in convention (SURVEY.md s8b A''): the host dlopen()s the library and resolves ONE C symbol
// (the reference: `extern "C" void trex_python_register()`, T/python/PythonEntryPoint.cpp:142-179), which fills a table of
// function pointers and registers the back ends -- detect::BackendHooks{init, deinit, is_initializing, fps, apply,
// set_background} (T/python/BackendRegistry.h:10-22; register_yolo_backend, T/python/YOLO.cpp:1738-1747) and
// RecTaskBackend{init, deinit, predict} (T/python/PythonBackendRegistry.cpp:52-58).  Like the reference's
// BackgroundSubtraction (static data(), T/python/BackgroundSubtraction.cpp:37-48) the table drives ONE process-wide detection
// handle and ONE identification handle; hosts that want several use the tb_seg_* / tb_vi_* functions directly.
// Host code only (no kernels); every entry forwards to the C ABI.
#include "common.h"

#include <chrono>
#include <mutex>
#include <vector>

namespace {

struct Backend {
    std::mutex mu;                       // the reference serialises apply() on the pipeline's manager mutex (TaskPipeline.h:226-259)
    tb_seg *seg = nullptr;
    tb_vi *vi = nullptr;
    bool has_bg = false;
    double fps_sum = 0.0, fps_samples = 0.0;      // BackgroundSubtraction::fps(): mean over apply() calls of frames / elapsed
    const tb_host_table *host = nullptr;
};
Backend &B() { static Backend b; return b; }

void log_host(int level, const char *msg)
{
    const tb_host_table *h = B().host;
    if (h && h->log) h->log(h->user, level, msg);
}

int be_init(const tb_seg_config *cfg, const tb_seg_params *params)
{
    std::lock_guard<std::mutex> g(B().mu);
    TB_REQUIRE(cfg, TB_ERR_INVALID, "tb_backend.init: null config");
    if (B().seg) { tb_seg_destroy(B().seg); B().seg = nullptr; B().has_bg = false; }
    int r = tb_seg_create(cfg, &B().seg);
    if (r != TB_OK) return r;
    if (params && (r = tb_seg_set_params(B().seg, params)) != TB_OK) { tb_seg_destroy(B().seg); B().seg = nullptr; return r; }
    B().fps_sum = B().fps_samples = 0.0;
    log_host(0, "trex_b200: background_subtraction back end initialised");
    return TB_OK;
}

void be_deinit(void)
{
    std::lock_guard<std::mutex> g(B().mu);
    if (B().seg) tb_seg_destroy(B().seg);
    B().seg = nullptr; B().has_bg = false;
}

// the reference's pipeline starts paused until a background is set (T/core/TaskPipeline.h:54-55,272-308)
int be_is_initializing(void)
{
    std::lock_guard<std::mutex> g(B().mu);
    return (B().seg && B().has_bg) ? 0 : 1;
}

double be_fps(void)
{
    std::lock_guard<std::mutex> g(B().mu);
    return B().fps_samples > 0 ? B().fps_sum / B().fps_samples : 0.0;
}

int be_set_background(const uint8_t *bg, int width, int height, int channels, int64_t stride)
{
    std::lock_guard<std::mutex> g(B().mu);
    TB_REQUIRE(B().seg, TB_ERR_STATE, "tb_backend.set_background: init() first");
    int r = tb_seg_set_background_c(B().seg, bg, width, height, channels, stride);
    if (r == TB_OK) B().has_bg = true;
    return r;
}

int be_update_params(const tb_seg_params *params)
{
    std::lock_guard<std::mutex> g(B().mu);
    TB_REQUIRE(B().seg, TB_ERR_STATE, "tb_backend.update_params: init() first");
    return tb_seg_set_params(B().seg, params);
}

int be_apply(const uint8_t *const *frames, int n, int64_t stride, tb_blob_view *views)
{
    std::lock_guard<std::mutex> g(B().mu);
    TB_REQUIRE(B().seg, TB_ERR_STATE, "tb_backend.apply: init() first");
    TB_REQUIRE(views, TB_ERR_INVALID, "tb_backend.apply: null views");
    const auto t0 = std::chrono::steady_clock::now();
    int r = tb_seg_submit(B().seg, frames, n, stride, 1);
    if (r != TB_OK) return r;
    r = tb_seg_wait(B().seg);
    if (r != TB_OK && r != TB_ERR_CAPACITY) return r;
    for (int i = 0; i < n; ++i) {
        const int q = tb_seg_result(B().seg, i, &views[i]);
        if (q != TB_OK) return q;
    }
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    B().fps_sum += n / std::max(dt, 1e-12);
    B().fps_samples += 1.0;
    return r;          // TB_ERR_CAPACITY: flagged frames carry tb_frame_info.status, the others are valid
}

int be_vi_init(const tb_vi_config *cfg)
{
    std::lock_guard<std::mutex> g(B().mu);
    TB_REQUIRE(cfg, TB_ERR_INVALID, "tb_backend.vi_init: null config");
    if (B().vi) { tb_vi_destroy(B().vi); B().vi = nullptr; }
    return tb_vi_create(cfg, &B().vi);
}

void be_vi_deinit(void)
{
    std::lock_guard<std::mutex> g(B().mu);
    if (B().vi) tb_vi_destroy(B().vi);
    B().vi = nullptr;
}

int be_vi_set_tensor(const char *name, const float *data, int64_t count)
{
    std::lock_guard<std::mutex> g(B().mu);
    TB_REQUIRE(B().vi, TB_ERR_STATE, "tb_backend.vi_set_tensor: vi_init() first");
    return tb_vi_set_tensor(B().vi, name, data, count);
}

int be_vi_commit(void)
{
    std::lock_guard<std::mutex> g(B().mu);
    TB_REQUIRE(B().vi, TB_ERR_STATE, "tb_backend.vi_commit: vi_init() first");
    return tb_vi_commit(B().vi);
}

int be_vi_predict(const uint8_t *images, int n, float *probs)
{
    std::lock_guard<std::mutex> g(B().mu);
    TB_REQUIRE(B().vi, TB_ERR_STATE, "tb_backend.vi_predict: vi_init() first");
    return tb_vi_predict(B().vi, images, n, probs, nullptr);
}

const tb_backend_table g_table = {
    TB_ABI_VERSION, (uint32_t)sizeof(tb_backend_table),
    be_init, be_deinit, be_is_initializing, be_fps, be_apply, be_set_background, be_update_params,
    be_vi_init, be_vi_deinit, be_vi_set_tensor, be_vi_commit, be_vi_predict, tb_last_error,
};

}  // namespace

extern "C" const tb_backend_table *tb_backend(void) { return &g_table; }

extern "C" int trex_b200_register(const tb_host_table *host)
{
    if (host) {
        TB_REQUIRE(host->abi_version == TB_ABI_VERSION && host->size >= sizeof(tb_host_table), TB_ERR_INVALID,
                   "trex_b200_register: the host was built against another TB_ABI_VERSION");
        B().host = host;
        if (host->register_backend) host->register_backend(host->user, "background_subtraction", &g_table);
    }
    return TB_OK;
}
