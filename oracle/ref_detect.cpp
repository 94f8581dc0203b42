// Test infrastructure (oracle/): a C ABI around the REFERENCE's own detection path -- the function BASELINE.json's north_star names:
//   tracker/python/BackgroundSubtraction.cpp          BackgroundSubtraction::set_background -> Data::set (:83-98), BackgroundSubtraction::apply(std::vector<TileImage>&&)
//                                                     (:126-347): colour conversion per meta_encoding, generate_binary, CPULabeling::run, the detect_size_filter test on
//                                                     pixels->size() * sqcm, the UINT16_MAX line rule, Frame::add_object
//   commons/common/processing/RawProcessing.cpp       RawProcessing::generate_binary (:262-683): invert, difference / absolute difference, blur_difference, threshold /
//                                                     inRange / adaptive threshold, closing, dilation / erosion with the re-threshold of the eroded mask, the final AND
//   commons/common/processing/{CPULabeling,Brototype,Source,DLList,ListCache}.cpp, tracker/core/SizeFilters.cpp, processing/Background.cpp (meta_encoding)
// all compiled unmodified into oracle/_ref/libref_detect.so (oracle/build_ref.py, -DREF_DETECT).  Every cv:: call of those files is forwarded to the REAL
// OpenCV (Python's cv2) through the callback the test installs (oracle/ref_stubs_detect/cv_detect.h).  The stand-ins around it (TileImage, pv::Frame as a
// collector, the pipeline registry, the settings table) are in oracle/ref_stubs_detect/.  Never linked into the product.
#include <python/BackgroundSubtraction.h>
#include <processing/RawProcessing.h>
#include <processing/Background.h>
#include <core/TrackingSettings.h>
#include <video/AveragingAccumulator.h>

using namespace cmn;

namespace track {
void Detection::apply_background_subtraction(std::vector<TileImage>&& tiles) { BackgroundSubtraction::apply(std::move(tiles)); }      // the friend of BackgroundSubtraction.h:20
}

static void fire(const char *name)
{
    for (auto &cb : ref_settings().cbs) cb(std::string_view(name));
}

extern "C" {

void ref_cv_set_bridge(cv::bridge_fn_t f) { cv::bridge() = f; }

// the bridge callback sizes its output with this and writes rows * cols * elemSize bytes to the returned pointer
void *ref_cv_out(void *mat, int rows, int cols, int type)
{
    auto *m = static_cast<cv::Mat *>(mat);
    m->create(rows, cols, type);
    return m->data;
}

// numeric / boolean settings by name (detect_threshold, threshold_maximum, use_closing, closing_size, dilation_size, image_invert, enable_difference,
// detect_threshold_is_absolute, use_adaptive_threshold, adaptive_threshold_scale, blur_difference, cm_per_pixel, tags_enable ...); the registered
// callbacks fire like GlobalSettings' do.  blur_difference is read ONCE, by the first generate_binary of a process (RawProcessing.cpp:267): the test
// loads a second copy of the library for the other value.
void ref_detect_setting(const char *name, double value)
{
    ref_settings().num[name] = value;
    fire(name);
}

void ref_detect_meta_encoding(int encoding)
{
    ref_settings().meta_encoding = encoding;
    (void)Background::meta_encoding();                 // registers Background.cpp's callbacks on first use
    fire("meta_encoding");
}

void ref_detect_size_filter(const double *ranges, int n)
{
    auto &f = detect_settings().size_filter;
    f.clear();
    for (int i = 0; i < n; ++i) f.emplace_back(ranges[2 * i], ranges[2 * i + 1]);
}

void ref_detect_color_channel(int c)
{
    if (c < 0) detect_settings().color_channel.reset();
    else detect_settings().color_channel = (uint8_t)c;
}

// AveragingAccumulator(method).add(frame) x n, finalize() (commons/common/video/AveragingAccumulator.cpp, compiled unmodified; cv::add / max / min / divide and the
// float -> 8-bit convertTo are the real OpenCV's through the bridge).  method: 0 mean, 1 mode, 2 max, 3 min; frames: n x rows x cols x ch; threaded != 0 uses add_threaded
int ref_average(const uint8_t *frames, int n, int rows, int cols, int ch, int method, int threaded, uint8_t *out)
{
    const averaging_method_t::Class m = method == 0 ? averaging_method_t::mean : (method == 1 ? averaging_method_t::mode : (method == 2 ? averaging_method_t::max : averaging_method_t::min));
    try {
        AveragingAccumulator acc(m);
        const size_t bytes = (size_t)rows * cols * ch;
        for (int i = 0; i < n; ++i) {
            cv::Mat f(rows, cols, (ch - 1) << 3);
            std::memcpy(f.data, frames + (size_t)i * bytes, bytes);
            if (threaded) acc.add_threaded(f); else acc.add(f);
        }
        auto img = acc.finalize();
        if ((size_t)img->rows * img->cols * img->dims != bytes) return -2;
        std::memcpy(out, img->data(), bytes);
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_average: %s\n", e.what());
        return -1;
    }
}

// RawProcessing(average).generate_binary(input, input, output): frame and average have `ch` (1 or 3) channels; out: rows * cols * ch bytes
int ref_generate_binary(const uint8_t *frame, int rows, int cols, int ch, const uint8_t *average, uint8_t *out)
{
    const int type = (ch - 1) << 3;
    cv::Mat avg(rows, cols, type), in(rows, cols, type), result;
    std::memcpy(avg.data, average, (size_t)rows * cols * ch);
    std::memcpy(in.data, frame, (size_t)rows * cols * ch);
    try {
        RawProcessing raw(avg, nullptr, nullptr);
        TagCache tags;
        raw.generate_binary(in, in, result, &tags);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_generate_binary: %s\n", e.what());
        return -1;
    }
    if (result.rows != rows || result.cols != cols || result.channels() != ch) return -2;
    for (int y = 0; y < rows; ++y) std::memcpy(out + (size_t)y * cols * ch, result.ptr(y), (size_t)cols * ch);
    return 0;
}

// BackgroundSubtraction::set_background(average) + BackgroundSubtraction::apply({one tile with one image}).  frame: rows x cols x frame_ch (3 or 4: the
// detection path receives colour frames); background: rows x cols x bg_ch (1 for gray / r3g3b2, 3 for rgb8).  Returns the number of objects the frame
// received, in the reference's order, packed like oracle/ref_labeling.cpp does; -1 if the promise carried an exception.
int64_t ref_background_subtraction_apply(const uint8_t *frame, int rows, int cols, int frame_ch, const uint8_t *background, int bg_ch,
                                         uint16_t *lines, int64_t cap_lines, uint8_t *pixels, int64_t cap_px, int64_t *line_off, int64_t *px_off, uint8_t *flags,
                                         int64_t cap_blobs, int32_t *encoding_out, int32_t *callbacks)
{
    auto bg = Image::Make((uint32_t)rows, (uint32_t)cols, (uint32_t)bg_ch);
    std::memcpy(bg->data(), background, (size_t)rows * cols * bg_ch);
    track::BackgroundSubtraction::set_background(std::move(bg));

    std::vector<TileImage> tiles(1);
    auto img = Image::Make((uint32_t)rows, (uint32_t)cols, (uint32_t)frame_ch);
    std::memcpy(img->data(), frame, (size_t)rows * cols * frame_ch);
    tiles[0].images.emplace_back(std::move(img));
    tiles[0].promise = std::make_unique<std::promise<SegmentationData>>();
    auto future = tiles[0].promise->get_future();
    int32_t called = 0;
    tiles[0].callback = [&called]() { ++called; };
    track::Detection::apply_background_subtraction(std::move(tiles));
    *callbacks = called;
    SegmentationData data;
    try {
        data = future.get();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_background_subtraction_apply: %s\n", e.what());
        return -1;
    }
    *encoding_out = (int32_t)data.frame.encoding().value();
    int64_t k = 0, nl = 0, np = 0;
    line_off[0] = 0; px_off[0] = 0;
    for (auto &b : data.frame.objects) {
        if (k >= cap_blobs) return -4;
        for (auto &h : *b.lines) {
            if (nl >= cap_lines) return -4;
            lines[4 * nl] = h.x0; lines[4 * nl + 1] = h.x1; lines[4 * nl + 2] = h.y; lines[4 * nl + 3] = 0; ++nl;
        }
        if (b.pixels) {
            if (np + (int64_t)b.pixels->size() > cap_px) return -4;
            std::memcpy(pixels + np, b.pixels->data(), b.pixels->size());
            np += (int64_t)b.pixels->size();
        }
        flags[k] = b.extra_flags;
        ++k;
        line_off[k] = nl; px_off[k] = np;
    }
    return k;
}

}
