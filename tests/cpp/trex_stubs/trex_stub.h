// Minimal mirror of the TRex declarations the INTEGRATION.md snippets touch -- TEST INFRASTRUCTURE (type-checks the snippets with
// `g++ -fsyntax-only`; the real headers pull in OpenCV / glaze / cnpy, which this image lacks: SURVEY.md s8c).  Every declaration
// cites the reference header it restates; tests/test_integration_snippets.py greps those headers (when /root/reference is
// present) so that a signature drifting in TRex fails the test instead of silently invalidating the snippets.
#pragma once
#include <cstdint>
#include <exception>
#include <functional>
#include <future>
#include <map>
#include <memory>
#include <mutex>
#include <optional>
#include <shared_mutex>
#include <stdexcept>
#include <string>
#include <vector>

#define TREX_EXPORT

namespace cmn {
struct Size2 { float width = 0, height = 0; };
struct Vec2 { float x = 0, y = 0; };
struct HorizontalLine { uint16_t x0, x1; uint16_t y, padding; };                 // C/misc/detail.h:71-131 (8 bytes)
using PixelArray_t = std::vector<uint8_t>;                                       // C/misc/types.h
// C/misc/Image.h:99-142
class Image {
public:
    using Ptr = std::unique_ptr<Image>;
    unsigned cols = 0, rows = 0, dims = 1;
    uint8_t *data() const { return _data; }
    size_t size() const { return (size_t)cols * rows * dims; }
    template <typename... Args> static Ptr Make(Args &&...) { return std::make_unique<Image>(); }
private:
    uint8_t *_data = nullptr;
};
namespace blob {
struct Prediction {};
using line_ptr_t = std::unique_ptr<std::vector<HorizontalLine>>;
using pixel_ptr_t = std::unique_ptr<PixelArray_t>;
struct Pair {                                                                    // C/misc/types.h:591-601
    line_ptr_t lines; pixel_ptr_t pixels; uint8_t extra_flags = 0; Prediction pred;
    Pair() = default;
    Pair(Pair &&) = default;
    Pair(line_ptr_t &&l, pixel_ptr_t &&p, uint8_t flags = 0, Prediction &&pr = {}) : lines(std::move(l)), pixels(std::move(p)), extra_flags(flags), pred(pr) {}
};
}
struct Timer { double elapsed() const { return 1.0; } };                         // C/misc/Timer.h
template <typename T, typename Construct, size_t N> struct ImageBuffers { void move_back(T &&) {} };   // C/misc/Buffers.h:186-260
struct SizeFilters { struct R { double start, end; }; std::vector<R> ranges() const { return {}; } };  // C/misc/SizeFilters.h
}
using namespace cmn;

enum class meta_encoding_t { gray, r3g3b2, rgb8, binary };                       // T/core/default_config.cpp (grab::default_config)
struct Background { static meta_encoding_t meta_encoding() { return meta_encoding_t::gray; } };   // C/processing/Background.h
inline uint8_t required_storage_channels(meta_encoding_t e) { return e == meta_encoding_t::rgb8 ? 3 : (e == meta_encoding_t::binary ? 0 : 1); }

namespace pv {
struct Frame {                                                                   // ProcessedVideo/pv.h:151-162
    void set_encoding(meta_encoding_t) {}
    void add_object(cmn::blob::Pair &&) {}
};
}
struct SegmentationData { pv::Frame frame; };                                    // T/core/DetectionImageTypes.h

struct TileImage {                                                               // T/core/TileImage.h:29-60
    Size2 tile_size;
    SegmentationData data;
    std::vector<Image::Ptr> images;
    std::unique_ptr<std::promise<SegmentationData>> promise;
    std::function<void()> callback;
};

namespace buffers {                                                              // T/core/TileBuffers.h:9-22
struct ImageMaker { cmn::Image::Ptr operator()() const { return cmn::Image::Make(); } };
struct TileBuffers {
    static constexpr size_t max_pool_size = 16;
    using Buffers_t = cmn::ImageBuffers<cmn::Image::Ptr, ImageMaker, max_pool_size>;
    static Buffers_t &get() { static Buffers_t b; return b; }
};
}

// settings access (C/misc/GlobalSettings.h: READ_SETTING(name, type); T/tracking/Tracker.h: FAST_SETTING(name))
template <typename T> T trex_setting(const char *) { return T{}; }
#define READ_SETTING(NAME, TYPE) trex_setting<TYPE>(#NAME)
#define FAST_SETTING(NAME) trex_setting<uint32_t>(#NAME)
namespace Settings { using cm_per_pixel_t = float; }
struct SoftException : std::runtime_error { using std::runtime_error::runtime_error; };   // C/misc/SoftException.h

template <typename T> struct PipelineManager { void set_paused(bool) {} bool is_terminated() const { return false; } };   // T/core/TaskPipeline.h:226-259

namespace track {
namespace detect {
struct ObjectDetectionType { enum Class { none, yolo, background_subtraction, precomputed }; };   // T/core/DetectionTypes.h
struct BackendHooks {                                                            // T/python/BackendRegistry.h:10-17
    std::function<void()> init;
    std::function<void()> deinit;
    std::function<bool()> is_initializing;
    std::function<double()> fps;
    std::function<void(std::vector<TileImage> &&)> apply;
    std::function<void(const cmn::Image::Ptr &)> set_background;
};
inline void register_backend(ObjectDetectionType::Class, BackendHooks) {}        // T/python/BackendRegistry.h:19
}
PipelineManager<TileImage> &manager();                                           // T/python/BackgroundSubtraction.cpp:46-48 (file static)

struct BackgroundSubtraction {                                                   // T/python/BackgroundSubtraction.h:10-27
    BackgroundSubtraction(cmn::Image::Ptr && = nullptr);
    static void set_background(cmn::Image::Ptr &&);
    static std::future<SegmentationData> apply(TileImage &&tiled);
    static void deinit();
    static double fps();
    static void apply(std::vector<TileImage> &&tiled);
    struct Data {                                                                // T/python/BackgroundSubtraction.cpp:12-44
        cmn::Image::Ptr _background;
        mutable std::shared_mutex _background_mutex;
        std::mutex _gpu_mutex;
        void set(cmn::Image::Ptr &&);
        void add_time_sample(double) {}
        double fps() const { return 0; }
    };
    static Data &data();
};
}

namespace Python {                                                               // T/ml/VisualIdentification.h:104-133
class VINetwork {
public:
    void load_weights_b200(const std::map<std::string, std::vector<float>> &state_dict);
    std::vector<float> probabilities(std::vector<cmn::Image::Ptr> &&images);
};
}
