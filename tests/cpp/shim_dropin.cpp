// N1 as far as this image allows: the drop-in shim of INTEGRATION.md s1 -- the replacement bodies of BackgroundSubtraction::Data::set and
// BackgroundSubtraction::apply(std::vector<TileImage>&&) -- compiled against the REFERENCE'S OWN tracker/python/BackgroundSubtraction.h (included from the
// checkout) and linked with libtrexb200.so, so that the reference's interface (set_background(Image::Ptr&&), TileImage with promise + callback, pv::Frame::add_object)
// drives the CUDA path.  The snippet itself is extracted from INTEGRATION.md by tests/build_dropin.py into integration_snippet_1.inc and included verbatim below.
// This file adds what a TRex maintainer KEEPS of BackgroundSubtraction.cpp around those two bodies (the Data record, the pipeline manager accessor, the
// forwarding statics), written here in the shortest form that satisfies the header; TileImage, Image, pv::Frame, the pipeline registry and the settings
// table are the test stand-ins of oracle/ref_stubs/ + oracle/ref_stubs_detect/ (TRex's own need OpenCV / glaze).  The C entry points come from
// oracle/ref_detect.cpp, compiled into the same library: tests/test_gpu_dropin_shim.py calls the same ref_background_subtraction_apply on this library and
// on the compiled reference (oracle/_ref/libref_detect.so) and compares what the two pv::Frames received.  Test infrastructure; never part of the product.
#include <python/BackgroundSubtraction.h>
#include <python/PipelineRegistry.h>
#include <processing/Background.h>
#include <core/TrackingSettings.h>
#include <core/TileBuffers.h>
#include <misc/Timer.h>

namespace track {

struct BackgroundSubtraction::Data {
    Image::Ptr _background;
    double _time{0.0}, _samples{0.0};
    mutable std::shared_mutex _time_mutex, _background_mutex, _gpu_mutex;
    void set(Image::Ptr&&);                                  // body: INTEGRATION.md s1
    double fps() { std::shared_lock g(_time_mutex); return _samples == 0 ? 0 : _time / _samples; }
    void add_time_sample(double s) { std::unique_lock g(_time_mutex); _time += s; _samples++; }
    bool has_background() const { std::shared_lock g(_background_mutex); return _background != nullptr; }
};

static PipelineManager<TileImage>& manager() { return detect::pipeline_manager(detect::ObjectDetectionType::background_subtraction); }

BackgroundSubtraction::BackgroundSubtraction(Image::Ptr&& average) { data().set(std::move(average)); }
void BackgroundSubtraction::set_background(Image::Ptr&& average) { data().set(std::move(average)); }
BackgroundSubtraction::Data& BackgroundSubtraction::data() { static Data d; return d; }
std::future<SegmentationData> BackgroundSubtraction::apply(TileImage&& tiled)
{
    tiled.promise = std::make_unique<std::promise<SegmentationData>>();
    auto f = tiled.promise->get_future();
    manager().enqueue(std::move(tiled));
    return f;
}
void BackgroundSubtraction::deinit() {}
double BackgroundSubtraction::fps() { return data().fps(); }

}

#include "_dropin/integration_snippet_1.inc"
