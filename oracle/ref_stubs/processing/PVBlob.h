// Test infrastructure: pv::Blob (commons/common/processing/PVBlob.h) reduced to what PixelTree.cpp and CPULabeling.cpp touch: the run list, the pixel
// bytes, the flag byte (Flags / set_flag / get_only_flag / copy_flags restated from PVBlob.h:138-168) and the bounding box of the runs
// (x = min x0, y = first y, width = max x1 - x + 1, height = last y - y + 1: what pv::Blob::init computes from the lines).
#ifdef REF_REAL_PVBLOB
#include_next <processing/PVBlob.h>
#else
#pragma once
#include <commons.pc.h>
#include <processing/Background.h>
#include <processing/BlobIdentity.h>
namespace pv {
// of CompressedBlob / ShortHorizontalLine (PVBlob.h:246-330) only what processing/BlobIdentity.cpp reads: the packed runs and the first row
struct ShortHorizontalLine {
    uint16_t _x0 = 0, _x1 = 0;
    constexpr ShortHorizontalLine() = default;
    constexpr ShortHorizontalLine(uint16_t x0, uint16_t x1, bool eol = false) : _x0(x0), _x1((x1 & 0x7FFF) | uint16_t(eol << 15)) {}
    constexpr uint16_t x0() const { return _x0; }
    constexpr uint16_t x1() const { return _x1 & 0x7FFF; }
};
struct CompressedBlob {
    uint16_t start_y{0};
    std::vector<ShortHorizontalLine> _lines;
    const std::vector<ShortHorizontalLine>& lines() const { return _lines; }
};
class Blob {
    cmn::blob::line_ptr_t _lines;
    cmn::blob::pixel_ptr_t _pixels;
    uint8_t _flags = 0;
    cmn::Bounds _bounds;
public:
    enum class Flags { split = 1, is_tag = 2, is_instance_segmentation = 4, is_rgb = 5, is_r3g3b2 = 6, is_binary = 7 };
    static constexpr void set_flag(uint8_t &flags, Flags flag, bool v) { flags ^= (-uint8_t(v) ^ flags) & (1UL << uint8_t(flag)); }
    static constexpr uint8_t get_only_flag(Flags flag, bool v) { uint8_t flags = 0; set_flag(flags, flag, v); return flags; }
    static constexpr bool is_flag(uint8_t flags, Flags flag) { return (flags >> uint8_t(flag)) & 1u; }
    static constexpr uint8_t flag(Flags flag) { return uint8_t(1UL << uint8_t(flag)); }
    Blob(cmn::blob::line_ptr_t&& l, cmn::blob::pixel_ptr_t&& p, uint8_t flags = 0, cmn::blob::Prediction&& = {}) : _lines(std::move(l)), _pixels(std::move(p)), _flags(flags)
    {
        if (_lines && !_lines->empty()) {
            int x0 = 1 << 30, x1 = -1;
            for (auto &h : *_lines) { x0 = std::min<int>(x0, h.x0); x1 = std::max<int>(x1, h.x1); }
            _bounds = cmn::Bounds((float)x0, (float)_lines->front().y, (float)(x1 - x0 + 1), (float)(_lines->back().y - _lines->front().y + 1));
        }
    }
    Blob(const Blob& o) : Blob(std::make_unique<cmn::blob::lines_t>(*o._lines), o._pixels ? std::make_unique<cmn::PixelArray_t>(*o._pixels) : nullptr, o._flags) {}
    template<typename... A> static BlobPtr Make(A&&... a) { return std::make_unique<Blob>(std::forward<A>(a)...); }
    const std::vector<cmn::HorizontalLine>& hor_lines() const { return *_lines; }
    const cmn::blob::line_ptr_t& lines() const { return _lines; }
    const cmn::blob::pixel_ptr_t& pixels() const { return _pixels; }
    cmn::blob::line_ptr_t&& steal_lines() { return std::move(_lines); }
    const cmn::Bounds& bounds() const { return _bounds; }
    // pv::Blob::add_offset (processing/PVBlob.cpp:1820-1837): every run moves by the integer part of the vector, clamped so that the box stays at >= 0
    void add_offset(const cmn::Vec2& off)
    {
        if (off == cmn::Vec2(0)) return;
        int offy = off.y, offx = off.x;
        if (!_lines->empty() && offy < -float(_bounds.y)) offy = -float(_bounds.y);
        if (!_lines->empty() && offx < -float(_bounds.x)) offx = -float(_bounds.x);
        for (auto &h : *_lines) { h.y += offy; h.x0 += offx; h.x1 += offx; }
        _bounds.x += (float)offx; _bounds.y += (float)offy;
    }
    // FilterCache.cpp's `moments` branch: the stand-in does not compute image moments -- the wrapper hands in the orientation (pv::Blob::calculate_moments,
    // PVBlob.cpp:111-213, stays restated-only in oracle/trex_oracle.c to_blob_orientation)
    float _orientation = 0;
    void calculate_moments() {}
    float orientation() const { return _orientation; }
    uint8_t flags() const { return _flags; }
    pv::bid blob_id() const { return pv::blob_bid(*this); }              // pv::Blob::init (PVBlob.cpp:785): the id is derived from the runs
    cmn::blob::Prediction prediction() const { return {}; }
    bool is_binary() const { return is_flag(_flags, Flags::is_binary); }
    bool is_rgb() const { return is_flag(_flags, Flags::is_rgb); }
    bool is_r3g3b2() const { return is_flag(_flags, Flags::is_r3g3b2); }
    uint8_t channels() const { return is_binary() ? 0 : (is_rgb() ? 3 : 1); }
    cmn::InputInfo input_info() const
    {
        cmn::InputInfo i{};
        i.channels = channels();
        i.encoding = is_rgb() ? cmn::meta_encoding_t::rgb8 : (is_r3g3b2() ? cmn::meta_encoding_t::r3g3b2 : (is_binary() ? cmn::meta_encoding_t::binary : cmn::meta_encoding_t::gray));
        return i;
    }
    size_t num_pixels() const { size_t n = 0; for (auto &h : *_lines) n += h.length(); return n; }
    static constexpr uint8_t copy_flags(const Blob& b)
    {
        return get_only_flag(Flags::is_rgb, b.is_rgb()) | get_only_flag(Flags::is_r3g3b2, b.is_r3g3b2()) | get_only_flag(Flags::is_binary, b.is_binary());
    }
};
}
#endif
