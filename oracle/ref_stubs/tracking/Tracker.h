// Test infrastructure: stands in for tracker/tracking/Tracker.h (quote-included by Outline.cpp and Posture.cpp).  Posture.cpp needs Tracker::background():
// the Background the test wrapper installs (oracle/ref_posture.cpp).  The pose-based overloads of calculate_posture in that file only have to COMPILE:
// blob::Pose, PoseMidlineIndexes, BasicStuff, SegmentedOutlines, the DLList cache, gui::reduce_vertex_line and cv::circle are inert look-alikes.
#pragma once
#include <commons.pc.h>
#include <processing/Background.h>
#include <processing/PVBlob.h>
#include <processing/DLList.h>
#include <misc/create_struct.h>
namespace cmn::blob {
struct Pose {
    struct Point { uint16_t x = 0, y = 0; bool valid() const { return x || y; } operator cmn::Vec2() const { return cmn::Vec2(x, y); } };
    std::vector<Point> points;
};
struct SegmentedOutlines { std::optional<std::vector<cmn::Vec2>> original_outline; };
}
namespace cmn::gui {
inline void reduce_vertex_line(const std::vector<cmn::Vec2>&, std::vector<cmn::Vec2>&, float) { std::fprintf(stderr, "reduce_vertex_line stand-in used\n"); std::abort(); }
}
namespace cv { template<typename... A> inline void circle(A&&...) { std::fprintf(stderr, "cv::circle stand-in used\n"); std::abort(); } }
namespace cmn {
struct ThreadSafePolicy {};
template<typename T, size_t N, template<typename...> class Ptr, typename Policy>
struct ObjectCache {
    Ptr<T> getObject() { return Ptr<T>(new T()); }
    void returnObject(Ptr<T>&&) {}
};
}
namespace track {
struct PoseMidlineIndexes { std::vector<uint8_t> indexes; };
struct BlobBoundsOnly { cmn::Bounds b; bool is_split = false; cmn::Bounds calculate_bounds() const { return b; } bool split() const { return is_split; } };
struct BasicStuff { BlobBoundsOnly blob; };
struct Tracker {
    static const cmn::Background *&background_slot() { static const cmn::Background *bg = nullptr; return bg; }
    static const cmn::Background *background() { return background_slot(); }
    struct Border { bool in_recognition_bounds(const cmn::Vec2&) const { return true; } };
    Border border() const { return Border{}; }
    static Tracker *instance() { static Tracker t; return &t; }
};
}
