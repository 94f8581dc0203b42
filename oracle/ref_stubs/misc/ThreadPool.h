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
// with an explicit thread count (misc/ThreadPool.h:126-195): N / threads items per package, the last package takes the remainder; run here one after the
// other on the calling thread -- the packages and their indices are the reference's, which is what pv::Blob::calculate_moments' per-thread sums depend on
template<typename F, typename I> inline void distribute_indexes(F&& fn, GenericThreadPool&, I start, I end, uint32_t threads)
{
    const int64_t N = std::distance(start, end);
    if (N <= 0) return;
    if (threads <= 1) { fn(int64_t(0), start, end, int64_t(0)); return; }
    const int64_t per_thread = std::max<int64_t>(1, N / int64_t(threads));
    int64_t i = 0, j = 0;
    I nex = start;
    for (auto it = start; it != end; ++j) {
        const int64_t step = (j + 1 == int64_t(threads)) ? N - i : per_thread;
        std::advance(nex, step);
        fn(i, it, nex, j);
        it = nex; i += step;
    }
}
}
