// Test infrastructure: the two templates of TRex's commons/common/misc/ranges.h that CircularGraph.cpp uses, restated (ranges.h:13-141, 149-200):
// Range<T>{start, end} with contains / length / iterable, and arange<T> whose k-th value is first + T(k) * step for k < size_t((last - first) / step).
#pragma once
#include <commons.pc.h>
namespace cmn {
template<typename T>
class arange {
public:
    T first, last, step;
    constexpr arange(T first = T(0), T last = T(0), T step = T(1)) : first(first), last(last), step(step) {}
    constexpr size_t num_steps() const { return size_t((last - first) / step); }
    struct iterator {
        const arange *p; size_t value;
        bool operator!=(const iterator& o) const { return value != o.value; }
        T operator*() const { return T(p->first + T(value) * T(p->step)); }
        iterator& operator++() { ++value; return *this; }
    };
    constexpr iterator begin() const { return iterator{this, 0}; }
    constexpr iterator end() const { return iterator{this, num_steps()}; }
};
template<typename T>
struct Range {
    T start, end;
    constexpr Range() noexcept : Range(T(), T()) {}
    explicit constexpr Range(T s, T e = T()) noexcept : start(s), end(e) {}
    constexpr bool empty() const { return start == end; }
    constexpr bool contains(T v) const { return v >= start && v < end; }
    constexpr bool operator<(const Range<T>& o) const { return start < o.start || (start == o.start && end < o.end); }
    constexpr T length() const { return end - start; }
    constexpr arange<T> iterable() const { return arange<T>(start, end); }
    constexpr bool operator==(const Range<T>& o) const { return o.start == start && o.end == end; }
};
using Rangel = Range<long_t>;
}
