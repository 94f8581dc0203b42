// Test infrastructure: look-alikes of what tracker/tracking/FilterCache.cpp's local_midline_length touches in tracker/tracking/Stuffs.h and Individual.h:
// PostureStuff (frame, outline size, cached midline length / angle), BasicStuff (tracking/Tracker.h of this directory), an Individual that is a list of
// frames with those two records, and Tracker::instance()->border().in_recognition_bounds (always inside).  The statistics themselves
// (Median<Float2_t>, standard_deviation, the step size) are the reference's code.
#pragma once
#include <commons.pc.h>
#include <misc/ranges.h>
#include <misc/Median.h>
#include <core/idx_t.h>
#include <tracking/Tracker.h>
namespace track {
struct TrackletInformation {};
struct OptionalFloat { float v = cmn::infinity<float>(); bool has_value() const { return v != cmn::infinity<float>(); } float value() const { return v; } };
struct OutlineSize { size_t n = 0; explicit operator bool() const { return n != 0; } size_t size() const { return n; } };
struct PostureStuff {
    cmn::Frame_t frame;
    OutlineSize outline;
    OptionalFloat midline_angle, midline_length;
    bool cached() const { return midline_length.has_value(); }
};
struct IdentityStandIn { Idx_t id; Idx_t ID() const { return id; } };
struct TrackletRange { cmn::Range<cmn::Frame_t> range; bool contains(cmn::Frame_t f) const { return f >= range.start && f <= range.end; } };
class Individual {
public:
    Idx_t id{0};
    std::vector<std::pair<BasicStuff, PostureStuff>> frames;       // consecutive frames from range.start
    cmn::Range<cmn::Frame_t> range;
    IdentityStandIn identity() const { return IdentityStandIn{id}; }
    TrackletRange get_tracklet(cmn::Frame_t) const { return TrackletRange{range}; }
    template<typename Fn> void iterate_frames(const cmn::Range<cmn::Frame_t>&, Fn&& fn, uint32_t step_size = 1u) const
    {
        const std::shared_ptr<TrackletInformation> none;
        for (size_t i = 0; i < frames.size(); i += step_size) {
            const BasicStuff *b = &frames[i].first; const PostureStuff *p = &frames[i].second;
            if (!fn(cmn::Frame_t(range.start.get() + (int32_t)i), none, b, p)) break;
        }
    }
};
}
