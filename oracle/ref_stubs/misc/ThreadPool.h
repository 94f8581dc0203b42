// Test infrastructure: stands in for commons/common/misc/ThreadPool.h.  Source::init only touches the pool when `enable_threads` is set; the tests run
// single-threaded, and distribute_indexes (should it be reached) runs the whole range as one chunk on the calling thread.
#pragma once
#include <commons.pc.h>
namespace cmn {
class GenericThreadPool {
public:
    GenericThreadPool(size_t, const std::string&) {}
};
template<typename F, typename I> inline void distribute_indexes(F&& fn, GenericThreadPool&, I start, I end) { fn(uint8_t(0), start, end, uint8_t(0)); }
}
