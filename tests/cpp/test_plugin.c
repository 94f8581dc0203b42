/* Pure-C host of the plug-in convention (SURVEY.md s8b A''): dlopen the library, resolve the ONE symbol trex_b200_register (the
 * reference resolves trex_python_register, T/python/PythonWrapper.cpp:402-470 / PythonEntryPoint.cpp:142-179), receive the
 * back-end table through the host's register_backend callback (detect::register_backend, T/python/BackendRegistry.h:19) and drive
 * detection + identification through the table alone.  Compiled as C (gcc -std=c11): the table carries no C++ types.
 *   test_plugin <path/to/libtrexb200.so> [gpu]
 * Without "gpu": registration and table shape only; init must fail with TB_ERR_CUDA when no device exists (no CPU fallback).
 * Prints "OK ..." on success. */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "trexb200.h"

static const tb_backend_table *g_table = NULL;
static char g_type[64];
static int g_logs = 0;

static void on_register(void *user, const char *detect_type, const tb_backend_table *t)
{
    *(int *)user += 1;
    strncpy(g_type, detect_type, sizeof(g_type) - 1);
    g_table = t;
}
static void on_log(void *user, int level, const char *msg) { (void)user; (void)level; (void)msg; ++g_logs; }

int main(int argc, char **argv)
{
    if (argc < 2) { printf("usage: test_plugin lib [gpu]\n"); return 2; }
    const int gpu = argc > 2 && !strcmp(argv[2], "gpu");
    void *so = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    if (!so) { printf("FAIL dlopen %s\n", dlerror()); return 1; }
    int (*reg)(const tb_host_table *) = (int (*)(const tb_host_table *))dlsym(so, "trex_b200_register");
    const tb_backend_table *(*get)(void) = (const tb_backend_table *(*)(void))dlsym(so, "tb_backend");
    if (!reg || !get) { printf("FAIL dlsym\n"); return 1; }
    int calls = 0;
    tb_host_table host = {TB_ABI_VERSION, (uint32_t)sizeof(tb_host_table), &calls, on_register, on_log};
    tb_host_table stale = host; stale.abi_version = TB_ABI_VERSION + 1;
    if (reg(&stale) != TB_ERR_INVALID || calls != 0) { printf("FAIL stale ABI accepted\n"); return 1; }
    if (reg(&host) != TB_OK || calls != 1 || strcmp(g_type, "background_subtraction") || g_table != get()) { printf("FAIL register\n"); return 1; }
    const tb_backend_table *t = g_table;
    if (t->abi_version != TB_ABI_VERSION || t->size != sizeof(tb_backend_table) || !t->init || !t->deinit || !t->is_initializing || !t->fps ||
        !t->apply || !t->set_background || !t->update_params || !t->vi_init || !t->vi_deinit || !t->vi_set_tensor || !t->vi_commit ||
        !t->vi_predict || !t->last_error) { printf("FAIL table shape\n"); return 1; }
    if (t->is_initializing() != 1 || t->fps() != 0.0) { printf("FAIL initial state\n"); return 1; }
    enum { W = 320, H = 240 };
    tb_seg_config cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.width = W; cfg.height = H; cfg.max_batch = 2; cfg.channels = 1;
    if (!gpu) {
        int rc = t->init(&cfg, NULL);
        if (rc == TB_OK) { t->deinit(); printf("OK registered (a device is present)\n"); return 0; }
        if (rc != TB_ERR_CUDA || !strstr(t->last_error(), "no CUDA device")) { printf("FAIL init without a device: %d %s\n", rc, t->last_error()); return 1; }
        printf("OK registered, init refused without a device\n");
        return 0;
    }
    tb_seg_params p;
    void (*defaults)(tb_seg_params *) = (void (*)(tb_seg_params *))dlsym(so, "tb_seg_default_params");
    defaults(&p);
    p.n_size_ranges = 1; p.size_lo[0] = 10; p.size_hi[0] = 100000;
    if (t->init(&cfg, &p) != TB_OK) { printf("FAIL init %s\n", t->last_error()); return 1; }
    if (t->is_initializing() != 1) { printf("FAIL paused until a background is set\n"); return 1; }
    uint8_t *bg = (uint8_t *)malloc(W * H), *fr = (uint8_t *)malloc(W * H);
    memset(bg, 100, W * H); memcpy(fr, bg, W * H);
    for (int x = 10; x < 30; ++x) fr[5 * W + x] = 20;      /* 20 px blob */
    fr[50 * W + 50] = 20;                                  /* 1 px: filtered */
    const uint8_t *frames[2] = {fr, bg};
    tb_blob_view v[2];
    if (t->apply(frames, 2, 0, v) != TB_ERR_STATE) { printf("FAIL apply before set_background\n"); return 1; }
    if (t->set_background(bg, W, H, 1, 0) != TB_OK || t->is_initializing() != 0) { printf("FAIL set_background %s\n", t->last_error()); return 1; }
    if (t->apply(frames, 2, 0, v) != TB_OK) { printf("FAIL apply %s\n", t->last_error()); return 1; }
    if (v[0].info.n_blobs != 1 || v[1].info.n_blobs != 0 || v[0].recs[0].n_pixels != 20 || v[0].lines[0].x0 != 10 || v[0].lines[0].x1 != 29 ||
        v[0].lines[0].y != 5 || v[0].pixels[0] != 20) { printf("FAIL blobs\n"); return 1; }
    p.detect_threshold = 90;                               /* |20 - 100| = 80 is no longer above the threshold */
    if (t->update_params(&p) != TB_OK || t->apply(frames, 2, 0, v) != TB_OK || v[0].info.n_blobs != 0) { printf("FAIL update_params\n"); return 1; }
    if (!(t->fps() > 0.0)) { printf("FAIL fps\n"); return 1; }
    /* identification through the same table: zero weights -> uniform probabilities */
    tb_vi_config vc; memset(&vc, 0, sizeof(vc));
    vc.width = 80; vc.height = 80; vc.channels = 1; vc.num_classes = 4; vc.max_images = 8; vc.precision = 1;
    if (t->vi_init(&vc) != TB_OK) { printf("FAIL vi_init %s\n", t->last_error()); return 1; }
    float probs[3 * 4];
    uint8_t *img = (uint8_t *)calloc(3, 6400);
    if (t->vi_predict(img, 3, probs) != TB_ERR_STATE) { printf("FAIL predict without weights\n"); return 1; }
    struct { const char *name; int n; float v; } ts[] = {
        {"model.conv1.weight", 16 * 25, 0.f}, {"model.conv1.bias", 16, 0.f}, {"model.conv2.weight", 64 * 16 * 25, 0.f}, {"model.conv2.bias", 64, 0.f},
        {"model.conv3.weight", 128 * 64 * 25, 0.f}, {"model.conv3.bias", 128, 0.f},
        {"model.bn1.weight", 16, 1.f}, {"model.bn1.bias", 16, 0.f}, {"model.bn1.running_mean", 16, 0.f}, {"model.bn1.running_var", 16, 1.f},
        {"model.bn2.weight", 64, 1.f}, {"model.bn2.bias", 64, 0.f}, {"model.bn2.running_mean", 64, 0.f}, {"model.bn2.running_var", 64, 1.f},
        {"model.bn3.weight", 128, 1.f}, {"model.bn3.bias", 128, 0.f}, {"model.bn3.running_mean", 128, 0.f}, {"model.bn3.running_var", 128, 1.f},
        {"model.fc1.weight", 100 * 12800, 0.f}, {"model.fc1.bias", 100, 0.f}, {"model.bn4.weight", 100, 1.f}, {"model.bn4.bias", 100, 0.f},
        {"model.fc2.weight", 4 * 100, 0.f}, {"model.fc2.bias", 4, 0.f}};
    for (size_t i = 0; i < sizeof(ts) / sizeof(ts[0]); ++i) {
        float *w = (float *)malloc(sizeof(float) * ts[i].n);
        for (int k = 0; k < ts[i].n; ++k) w[k] = ts[i].v;
        if (t->vi_set_tensor(ts[i].name, w, ts[i].n) != TB_OK) { printf("FAIL vi_set_tensor\n"); return 1; }
        free(w);
    }
    if (t->vi_commit() != TB_OK || t->vi_predict(img, 3, probs) != TB_OK) { printf("FAIL vi %s\n", t->last_error()); return 1; }
    for (int i = 0; i < 12; ++i) if (probs[i] < 0.2499f || probs[i] > 0.2501f) { printf("FAIL probs[%d]=%f\n", i, probs[i]); return 1; }
    t->vi_deinit(); t->deinit();
    if (t->is_initializing() != 1) { printf("FAIL deinit\n"); return 1; }
    printf("OK plug-in: %d log lines\n", g_logs);
    free(bg); free(fr); free(img);
    dlclose(so);
    return 0;
}
