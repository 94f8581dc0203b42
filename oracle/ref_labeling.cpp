// Test infrastructure (oracle/): a C ABI around the REFERENCE's own run-length connected-component labeling -- commons/common/processing/
// {CPULabeling,Brototype,Source,DLList,ListCache}.cpp (CPULabeling::run: Source::extract_lines -> merge_lines with Brototype -> run_fast), compiled
// unmodified from the reference checkout (oracle/build_ref.py; stand-ins in oracle/ref_stubs/, cv::Mat as a plain byte image).  Called like
// RawProcessing / BackgroundSubtraction::apply (run(image, cache)) and like pixel::threshold_blob (run(lines, pixels, cache, channels)) call it.
// Never linked into the product.
#include <processing/CPULabeling.h>
#include <processing/PVBlob.h>

using namespace cmn;

static int64_t store(blobs_t& blobs, uint16_t *lines, int64_t cap_lines, uint8_t *pixels, int64_t cap_px, int64_t *line_off, int64_t *px_off, uint8_t *flags, int64_t cap_blobs)
{
    int64_t k = 0, nl = 0, np = 0;
    line_off[0] = 0; px_off[0] = 0;
    for (auto &b : blobs) {
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

extern "C" {

// pv::blob_bid (processing/BlobIdentity.cpp:7-16 over pv::bid::from_data, misc/bid.h:87-94 -- both the reference's own files): the id pv::Blob::init
// gives a blob, from its first run and its number of runs (the count passes through from_data's uint8_t parameter)
uint32_t ref_blob_bid(const uint16_t *in_lines, int64_t n)
{
    auto l = std::make_unique<blob::lines_t>((size_t)n);
    for (int64_t i = 0; i < n; ++i) (*l)[(size_t)i] = HorizontalLine(in_lines[4 * i + 2], in_lines[4 * i], in_lines[4 * i + 1]);
    pv::Blob blob(std::move(l), nullptr, 0);
    return (uint32_t)pv::blob_bid(blob);
}

// CPULabeling::run(image, cache): every non-zero pixel (any channel for 3-channel images) is foreground.  Blobs in the reference's emission order.
int64_t ref_label_image(const uint8_t *img, int rows, int cols, int channels, uint16_t *lines, int64_t cap_lines, uint8_t *pixels, int64_t cap_px,
                        int64_t *line_off, int64_t *px_off, uint8_t *flags, int64_t cap_blobs)
{
    cv::Mat m(rows, cols, channels == 3 ? CV_8UC3 : CV_8UC1);
    std::memcpy(m.data, img, (size_t)rows * cols * channels);
    CPULabeling::ListCache_t cache;
    auto blobs = CPULabeling::run(m, cache, false);
    return store(blobs, lines, cap_lines, pixels, cap_px, line_off, px_off, flags, cap_blobs);
}

// CPULabeling::run(lines, pixels, cache, channels): the entry pixel::threshold_blob uses on the runs that survive the tracker's threshold
int64_t ref_label_lines(const uint16_t *in_lines, int64_t n, const uint8_t *in_px, int64_t n_px, int channels, uint16_t *lines, int64_t cap_lines, uint8_t *pixels,
                        int64_t cap_px, int64_t *line_off, int64_t *px_off, uint8_t *flags, int64_t cap_blobs)
{
    std::vector<HorizontalLine> l((size_t)n);
    for (int64_t i = 0; i < n; ++i) l[(size_t)i] = HorizontalLine(in_lines[4 * i + 2], in_lines[4 * i], in_lines[4 * i + 1]);
    std::vector<uchar> px(in_px, in_px + n_px);
    CPULabeling::ListCache_t cache;
    auto blobs = CPULabeling::run(l, std::span<uchar>(px.data(), px.size()), cache, (uint8_t)channels);
    return store(blobs, lines, cap_lines, pixels, cap_px, line_off, px_off, flags, cap_blobs);
}

}
