// Test infrastructure: tracker/core/TileBuffers.h reduced to the call BackgroundSubtraction::apply makes when it hands a tile's images back to the pool.
#pragma once
#include <commons.pc.h>
namespace buffers {
struct TileBuffers {
    struct Buffers_t { size_t returned = 0; void move_back(cmn::Image::Ptr&& p) { p.reset(); ++returned; } };
    static Buffers_t& get() { static Buffers_t b; return b; }
};
}
