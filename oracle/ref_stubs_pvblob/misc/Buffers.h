// Test infrastructure: commons/common/misc/Buffers.h reduced to the recycling pool of run lists PVBlob.cpp keeps (pv::buffers(): get / move_back).
#pragma once
#include <commons.pc.h>
#include <source_location>
namespace cmn {
using source_location = std::source_location;
template<typename T, typename Construct, size_t N = 0>
class Buffers {
    std::vector<T> _pool; Construct _create{};
public:
    T get(const source_location&) { if (_pool.empty()) return _create(); T t = std::move(_pool.back()); _pool.pop_back(); return t; }
    void move_back(T&& t) { if (t) _pool.emplace_back(std::move(t)); }
    size_t size() const { return _pool.size(); }
};
}
