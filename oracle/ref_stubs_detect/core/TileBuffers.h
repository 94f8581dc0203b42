// Test infrastructure: tracker/core/TileBuffers.h reduced to the call BackgroundSubtraction::apply makes when it hands a tile's images back to the pool.
#pragma once
#include <commons.pc.h>
namespace buffers {
struct TileBuffers {
    struct Buffers_t { std::vector<cmn::Image::Ptr> pool; void move_back(cmn::Image::Ptr&& p) { pool.emplace_back(std::move(p)); } };      // a pool keeps its buffers alive (the shim page-locks them)
    static Buffers_t& get() { static Buffers_t b; return b; }
};
}
