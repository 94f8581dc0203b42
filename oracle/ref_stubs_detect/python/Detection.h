// Test infrastructure (oracle/): look-alikes of what tracker/python/BackgroundSubtraction.{h,cpp} touch around the function under test
// (BackgroundSubtraction::apply(std::vector<TileImage>&&)): the detection type enumeration, TileImage / SegmentationData (tracker/core/TileImage.h,
// TaskPipeline.h:87-118) with a pv::Frame that COLLECTS what add_object receives (ProcessedVideo/pv.cpp:491-529 appends the pair; the flags it sets follow
// from the encoding), a pipeline manager that never runs a thread, the ObjectDetection concept (tracker/python/Detection.h:13-24), and `struct Detection`,
// the friend BackgroundSubtraction.h names -- which is how the test wrapper reaches the private multi-tile apply().  Never linked into the product.
#pragma once
#include <commons.pc.h>
#include <future>
#include <processing/encoding.h>
#include <processing/PVBlob.h>
#ifndef TREX_EXPORT
#define TREX_EXPORT
#endif
namespace pv {
class Frame {
public:
    cmn::meta_encoding_t::Class _encoding = cmn::meta_encoding_t::gray;
    std::vector<cmn::blob::Pair> objects;
    void set_encoding(cmn::meta_encoding_t::Class e) { _encoding = e; }
    cmn::meta_encoding_t::Class encoding() const { return _encoding; }
    void add_object(cmn::blob::Pair&& pair)              // pv.cpp:491-529: empty objects are refused, the encoding sets three flags, the pair is appended
    {
        if (pair.lines->empty()) return;
        Blob::set_flag(pair.extra_flags, Blob::Flags::is_rgb, _encoding == cmn::meta_encoding_t::rgb8);
        Blob::set_flag(pair.extra_flags, Blob::Flags::is_r3g3b2, _encoding == cmn::meta_encoding_t::r3g3b2);
        Blob::set_flag(pair.extra_flags, Blob::Flags::is_binary, _encoding == cmn::meta_encoding_t::binary);
        objects.emplace_back(std::move(pair));
    }
};
}
struct SegmentationData {
    cmn::Image::Ptr image;
    pv::Frame frame;
};
struct TileImage {
    cmn::Size2 tile_size;
    SegmentationData data;
    std::vector<cmn::Image::Ptr> images;
    std::unique_ptr<std::promise<SegmentationData>> promise;
    std::function<void()> callback;
    TileImage() = default;
    TileImage(TileImage&&) = default;
    TileImage& operator=(TileImage&&) = default;
};
template<typename T>
class PipelineManager {
public:
    bool paused = true;
    std::vector<T> queue;
    bool is_terminated() const { return false; }
    void set_paused(bool p) { paused = p; }
    void enqueue(T&& t) { queue.emplace_back(std::move(t)); }
};
namespace track {
namespace detect {
namespace ObjectDetectionType { enum Class { none, yolo, background_subtraction, precomputed }; }
inline PipelineManager<TileImage>& pipeline_manager(ObjectDetectionType::Class) { static PipelineManager<TileImage> m; return m; }
inline void register_pipeline(ObjectDetectionType::Class, size_t, bool, std::function<void(std::vector<TileImage>&&)>) {}
inline void unregister_pipeline(ObjectDetectionType::Class) {}
}
template<typename T> concept MultiObjectDetection = requires (std::vector<TileImage> tiles) { { T::apply(std::move(tiles)) }; };
template<typename T> concept SingleObjectDetection = requires (TileImage tiles) { { T::apply(std::move(tiles)) } -> std::convertible_to<std::future<SegmentationData>>; };
template<typename T> concept ObjectDetection = MultiObjectDetection<T> || SingleObjectDetection<T>;
struct BackgroundSubtraction;
struct Detection { static void apply_background_subtraction(std::vector<TileImage>&& tiles); };      // defined in oracle/ref_detect.cpp
}
