// Test infrastructure: the basic types TRex's commons.pc.h pulls in through misc/types.h / misc/detail.h, as far as the compiled reference files use
// them.  The containers are the REFERENCE'S OWN (misc/IllegalVector.h is included from the checkout); the rest restates
//   coord_t = uint16_t, ptr_safe_t = uint64_t, HorizontalLine {x0, x1, y, padding}   commons/common/misc/detail.h:71-116
//   PixelArray_t = IllegalArray<uchar>, blob::lines_t / line_ptr_t / pixel_ptr_t / Pair  commons/common/misc/types.h:101,206-209,343-370
//   contains / insert_sorted / find_sorted                                             commons/common/commons.pc.h:640-716
#pragma once
#include <deque>
#include <span>
#include <thread>
#include <unordered_set>
using uchar = unsigned char;
#include <misc/IllegalVector.h>

namespace cmn {
using coord_t = uint16_t;
using ptr_safe_t = uint64_t;
template<typename T> class NoInitializeAllocator : public std::allocator<T> {      // misc/types.h:63-95 (skips default construction; std::vector<T> copies from it as a plain allocator)
public:
    template<typename U> struct rebind { typedef NoInitializeAllocator<U> other; };
    NoInitializeAllocator() noexcept : std::allocator<T>() {}
    NoInitializeAllocator(const NoInitializeAllocator<T>& rhs) noexcept : std::allocator<T>(rhs) {}
};
struct HorizontalLine {
    coord_t x0, x1;
    coord_t y, padding;
    constexpr HorizontalLine() noexcept = default;
    constexpr HorizontalLine(coord_t y_, coord_t x0_, coord_t x1_) noexcept : x0(x0_), x1(x1_), y(y_), padding(0) {}
    constexpr bool overlap_x(const HorizontalLine& o) const noexcept { return o.x1 >= x0 - 1 && o.x0 <= x1 + 1; }
    constexpr bool operator==(const HorizontalLine& o) const noexcept { return o.x0 == x0 && o.y == y && o.x1 == x1; }
    constexpr bool operator<(const HorizontalLine& o) const noexcept { return y < o.y || (y == o.y && x0 < o.x0); }
    constexpr ptr_safe_t length() const noexcept { return ptr_safe_t(x1) - ptr_safe_t(x0) + 1; }
    // pv::Blob::init's repair of unordered run lists (misc/detail.cpp): never reached by ordered input
    template<typename... A> static void repair_lines_array(A&...) { std::fprintf(stderr, "HorizontalLine::repair_lines_array stand-in used\n"); std::abort(); }
};
using PixelArray_t = IllegalArray<uchar>;
template<typename T, typename... A> constexpr bool is_in(const T& v, const A&... a) { return ((v == T(a)) || ...); }
template<typename Cont, typename V> inline bool contains(const Cont& c, const V& v) { return std::find(c.begin(), c.end(), v) != c.end(); }
template<class T, typename Cmp> inline auto insert_sorted(std::vector<T>& v, T&& e, Cmp&& cmp) { return v.insert(std::upper_bound(v.begin(), v.end(), e, std::forward<Cmp>(cmp)), std::move(e)); }
template<class T, typename Cmp, class K = T> inline auto find_sorted(const std::vector<T>& v, const K& e, Cmp&& cmp)
{
    auto it = std::lower_bound(v.begin(), v.end(), e, cmp);
    return (it != v.end() && !cmp(e, *it)) ? it : v.end();
}
inline unsigned hardware_concurrency() { return 1u; }

namespace blob {
using lines_t = std::vector<HorizontalLine>;
using line_ptr_t = std::unique_ptr<lines_t>;
using pixel_ptr_t = std::unique_ptr<PixelArray_t>;
struct Prediction {};
struct Pair {
    line_ptr_t lines; pixel_ptr_t pixels; uint8_t extra_flags = 0; Prediction pred;
    Pair() = default;
    Pair(line_ptr_t&& l, pixel_ptr_t&& p, uint8_t f = 0, Prediction&& pr = {}) : lines(std::move(l)), pixels(std::move(p)), extra_flags(f), pred(pr) {}
};
}
using blobs_t = std::vector<blob::Pair>;
}
namespace pv { class Blob; using BlobPtr = std::unique_ptr<Blob>; }
namespace cmn::blob { struct Pose; struct SegmentedOutlines; }      // named by tracker/tracking/Posture.h; look-alikes in tracking/Tracker.h of this directory
