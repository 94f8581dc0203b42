// Drives the C++ host layer (include/trexb200.hpp) the way TRex's test_matching.cpp:1556-1602 drives
// CPULabeling::run: circle + rectangle -> exactly one blob; render -> relabel -> identical lines.
// Prints "OK <n_lines> <n_pixels>" on success. Built and run by tests/test_gpu_cpp_shim.py.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "trexb200.hpp"

int main()
{
    const int W = 320, H = 240;
    std::vector<uint8_t> img((size_t)W * H, 0);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            if ((y - 100) * (y - 100) + (x - 120) * (x - 120) <= 50 * 50) img[(size_t)y * W + x] = 255;
            if (y >= 90 && y < 160 && x >= 150 && x < 260) img[(size_t)y * W + x] = 200;
        }
    try {
        auto blobs = trexb200::labeling_run(img.data(), W, H);
        if (blobs.size() != 1) { std::printf("FAIL blobs=%zu\n", blobs.size()); return 1; }
        std::vector<uint8_t> render((size_t)W * H, 0);
        size_t o = 0;
        for (auto &l : *blobs[0].lines)
            for (int x = l.x0; x <= l.x1; ++x) render[(size_t)l.y * W + x] = (*blobs[0].pixels)[o++];
        if (render != img) { std::printf("FAIL render differs\n"); return 1; }
        auto again = trexb200::labeling_run(render.data(), W, H);
        if (again.size() != 1 || again[0].lines->size() != blobs[0].lines->size() || *again[0].pixels != *blobs[0].pixels) {
            std::printf("FAIL not idempotent\n"); return 1;
        }
        // threshold path with a background + size filter, as BackgroundSubtraction::apply
        trexb200::BackgroundSubtraction bs(W, H, 2);
        std::vector<uint8_t> bg((size_t)W * H, 100), fr(bg);
        for (int x = 10; x < 30; ++x) fr[(size_t)5 * W + x] = 20;      // 20 px blob
        fr[(size_t)50 * W + 50] = 20;                                  // 1 px blob: filtered
        bs.settings().n_size_ranges = 1; bs.settings().size_lo[0] = 10; bs.settings().size_hi[0] = 100000;
        bs.update_settings();
        bs.set_background(bg.data());
        auto res = bs.apply({fr.data(), bg.data()});
        if (res[0].size() != 1 || res[1].size() != 0 || res[0][0].pixels->size() != 20) { std::printf("FAIL apply\n"); return 1; }
        // outline of the 20 x 1 blob: 42 side midpoints, blob on the right hand, no resampling
        auto ol = bs.outlines(0.f);
        if (ol.size() != 1 || ol[0].size() != 2 * 42 || ol[0][0] != 0.0f || ol[0][1] != 0.5f) { std::printf("FAIL outlines %zu\n", ol.empty() ? (size_t)0 : ol[0].size()); return 1; }
        // its midline: tail at index 0 of the walked outline, head on the far side, more than two segments along the bar
        auto ml = bs.midlines(1.f);
        if (ml.size() != 1 || ml[0].tail_index != 0 || ml[0].head_index <= 0 || ml[0].segments.size() < 4 * 3) { std::printf("FAIL midlines\n"); return 1; }
        // the whole posture chain (Individual::calculate_midline_for): a normalised midline of midline_resolution = 25 segments whose length
        // is about the bar's, first point at the origin; and pv::Blob::recount of the 20 px blob (|100 - 20| = 80 >= threshold)
        auto nm = bs.posture(1.f);
        if (nm.size() != 1 || nm[0].segments.size() != 4 * 25 || nm[0].segments[0] != 0.f || nm[0].segments[1] != 0.f || nm[0].len < 10.f || nm[0].len > 22.f) {
            std::printf("FAIL posture %zu %f\n", nm.empty() ? (size_t)0 : nm[0].segments.size(), nm.empty() ? 0.f : nm[0].len); return 1;
        }
        auto rc = bs.recount(50), rc2 = bs.recount(90);
        if (rc.size() != 1 || rc[0] != 20.f || rc2[0] != 0.f) { std::printf("FAIL recount\n"); return 1; }
        // colour frames: BGRA input, meta_encoding rgb8 -> B,G,R per blob pixel; gray frames are refused for rgb8
        {
            trexb200::BackgroundSubtraction cs(W, H, 1, 0, 0, 4, trexb200::meta_encoding_t::rgb8);
            std::vector<uint8_t> bg3((size_t)W * H * 3, 100), f4((size_t)W * H * 4, 100);
            for (int x = 10; x < 30; ++x) { uint8_t *q = &f4[((size_t)5 * W + x) * 4]; q[0] = 10; q[1] = 20; q[2] = 30; }
            cs.settings().n_size_ranges = 0;
            cs.update_settings();
            cs.set_background(bg3.data());
            auto cr = cs.apply({f4.data()});
            if (cr[0].size() != 1 || cr[0][0].pixels->size() != 60 || (*cr[0][0].pixels)[0] != 10 || (*cr[0][0].pixels)[2] != 30 ||
                cr[0][0].extra_flags == 0) { std::printf("FAIL rgb8\n"); return 1; }
            bool refused = false;
            try { trexb200::BackgroundSubtraction bad(W, H, 1, 0, 0, 1, trexb200::meta_encoding_t::rgb8); } catch (const std::exception &) { refused = true; }
            if (!refused) { std::printf("FAIL rgb8 accepted gray frames\n"); return 1; }
        }
        // identification through the C++ wrapper: set_tensor / commit / probabilities / paverages (VisualIdentification.h:92-181)
        {
            const int M = 5;
            trexb200::VINetwork net(M, 16);
            bool threw = false;
            std::vector<uint8_t> c0(6400, 0), c1(6400, 0), c2(6400, 0);
            for (int y = 20; y < 60; ++y) for (int x = 10; x < 70; ++x) { c0[(size_t)y * 80 + x] = (uint8_t)(x + y); c1[(size_t)y * 80 + x] = (uint8_t)(3 * x); }
            for (int i = 0; i < 6400; ++i) c2[(size_t)i] = (uint8_t)(i * 7);
            try { net.probabilities({c0.data()}); } catch (const std::exception &) { threw = true; }        // no weights: SoftException in the reference
            if (!threw) { std::printf("FAIL probabilities without weights\n"); return 1; }
            uint32_t seed = 12345u;
            auto fill = [&](size_t n, float scale, float offset) {
                std::vector<float> v(n);
                for (auto &x : v) { seed = seed * 1664525u + 1013904223u; x = offset + scale * ((float)(seed >> 8) / 16777216.0f - 0.5f); }
                return v;
            };
            const struct { const char *name; size_t n; float scale, offset; } ts[] = {
                {"model.conv1.weight", 16 * 25, 0.006f, 0.f}, {"model.conv1.bias", 16, 0.1f, 0.f}, {"model.conv2.weight", 64 * 16 * 25, 0.1f, 0.f}, {"model.conv2.bias", 64, 0.1f, 0.f},
                {"model.conv3.weight", 128 * 64 * 25, 0.05f, 0.f}, {"model.conv3.bias", 128, 0.1f, 0.f},
                {"model.bn1.weight", 16, 0.5f, 1.f}, {"model.bn1.bias", 16, 0.2f, 0.f}, {"model.bn1.running_mean", 16, 0.2f, 0.f}, {"model.bn1.running_var", 16, 0.5f, 1.f},
                {"model.bn2.weight", 64, 0.5f, 1.f}, {"model.bn2.bias", 64, 0.2f, 0.f}, {"model.bn2.running_mean", 64, 0.2f, 0.f}, {"model.bn2.running_var", 64, 0.5f, 1.f},
                {"model.bn3.weight", 128, 0.5f, 1.f}, {"model.bn3.bias", 128, 0.2f, 0.f}, {"model.bn3.running_mean", 128, 0.2f, 0.f}, {"model.bn3.running_var", 128, 0.5f, 1.f},
                {"model.fc1.weight", 100 * 12800, 0.02f, 0.f}, {"model.fc1.bias", 100, 0.1f, 0.f}, {"model.bn4.weight", 100, 0.5f, 1.f}, {"model.bn4.bias", 100, 0.2f, 0.f},
                {"model.fc2.weight", (size_t)M * 100, 0.4f, 0.f}, {"model.fc2.bias", (size_t)M, 0.2f, 0.f}};
            for (auto &t : ts) net.set_tensor(t.name, fill(t.n, t.scale, t.offset));
            net.commit();
            auto pr = net.probabilities({c0.data(), c1.data(), c2.data(), c0.data()});
            if (pr.size() != 4 * (size_t)M) { std::printf("FAIL probabilities size\n"); return 1; }
            for (int i = 0; i < 4; ++i) {
                float sum = 0; for (int k = 0; k < M; ++k) { sum += pr[(size_t)i * M + k]; if (!(pr[(size_t)i * M + k] > 0.f)) { std::printf("FAIL prob <= 0\n"); return 1; } }
                if (sum < 0.9999f || sum > 1.0001f) { std::printf("FAIL softmax row sums to %f\n", sum); return 1; }
            }
            bool differ = false;
            for (int k = 0; k < M; ++k) { if (pr[(size_t)k] != pr[(size_t)3 * M + k]) { std::printf("FAIL same crop, different row\n"); return 1; } differ |= pr[(size_t)k] != pr[(size_t)M + k]; }
            if (!differ) { std::printf("FAIL different crops, same row\n"); return 1; }
            auto av = net.paverages({7u, 9u, 7u, 7u}, {c0.data(), c1.data(), c2.data(), c0.data()});
            if (av.size() != 2 || av[7u].samples != 3 || av[9u].samples != 1) { std::printf("FAIL paverages groups\n"); return 1; }
            for (int k = 0; k < M; ++k) {
                const float e7 = ((pr[(size_t)k] + 0.f) + pr[(size_t)2 * M + k] + pr[(size_t)3 * M + k]) / 3.f;
                if (av[9u].values[(size_t)k] != pr[(size_t)M + k] || std::fabs(av[7u].values[(size_t)k] - e7) > 1e-6f) { std::printf("FAIL paverages values\n"); return 1; }
            }
            trexb200::HostFrame pinned((size_t)W * H);         // page-locked frame buffer of the pool: same results as from pageable memory
            std::memcpy(pinned.data(), fr.data(), (size_t)W * H);
            auto again2 = bs.apply({pinned.data()});
            if (again2[0].size() != 1 || again2[0][0].pixels->size() != 20) { std::printf("FAIL pinned apply\n"); return 1; }
        }
        std::printf("OK %zu %zu\n", blobs[0].lines->size(), blobs[0].pixels->size());
    } catch (const std::exception &e) {
        std::printf("EXC %s\n", e.what());
        return 2;
    }
    return 0;
}
