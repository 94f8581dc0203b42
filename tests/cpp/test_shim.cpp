// Drives the C++ host layer (include/trexb200.hpp) the way TRex's test_matching.cpp:1556-1602 drives
// CPULabeling::run: circle + rectangle -> exactly one blob; render -> relabel -> identical lines.
// Prints "OK <n_lines> <n_pixels>" on success. Built and run by tests/test_gpu_cpp_shim.py.
#include <cstdio>
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
        std::printf("OK %zu %zu\n", blobs[0].lines->size(), blobs[0].pixels->size());
    } catch (const std::exception &e) {
        std::printf("EXC %s\n", e.what());
        return 2;
    }
    return 0;
}
