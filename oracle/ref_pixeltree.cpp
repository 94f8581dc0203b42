// Test infrastructure (oracle/): a C ABI around the REFERENCE's own pixel::find_outer_points (commons/common/processing/PixelTree.cpp:497-651 with
// pixel::Tree::add / generate_edges / walk, :652-1133), compiled unmodified from the reference checkout (oracle/build_ref.py; stand-ins in
// oracle/ref_stubs/processing/pixeltree_standins.h).  Called like posture::calculate_posture does (tracker/tracking/Posture.cpp:330-348).
// Never linked into the product.
#include <processing/PixelTree.h>
#include <processing/PVBlob.h>

extern "C" {

// lines: n x {x0, x1, y, pad} (the memory layout of HorizontalLine).  pts: all outlines back to back, off[k] .. off[k + 1] = outline k.
// Returns the number of outlines, -4 when a capacity is too small.
int64_t ref_find_outer_points(const uint16_t *lines, int64_t n, float *pts, int64_t cap_pts, int64_t *off, int64_t cap_off)
{
    auto l = std::make_unique<cmn::blob::lines_t>((size_t)n);
    for (int64_t i = 0; i < n; ++i) (*l)[(size_t)i] = cmn::HorizontalLine(lines[4 * i + 2], lines[4 * i], lines[4 * i + 1]);
    pv::Blob blob(std::move(l), nullptr);
    auto outlines = cmn::pixel::find_outer_points(&blob, 0);
    int64_t k = 0, total = 0;
    off[0] = 0;
    for (auto &o : outlines) {
        if (k + 1 > cap_off || total + (int64_t)o->size() > cap_pts) return -4;
        for (auto &p : *o) { pts[2 * total] = p.x; pts[2 * total + 1] = p.y; ++total; }
        off[++k] = total;
    }
    return k;
}

}
