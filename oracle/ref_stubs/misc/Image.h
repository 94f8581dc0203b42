// Test infrastructure: placeholder for commons/common/misc/Image.h; the cmn::Image stand-in (rows, cols, dims, data(), get() as a cv::Mat view) lives in misc/detail_bits.h.
#pragma once
#include <commons.pc.h>
