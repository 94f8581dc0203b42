// Test infrastructure: brings in the REFERENCE'S OWN commons/common/misc/bid.h (pv::bid: from_data, xy2d / d2xy, calc_position) from the checkout; this file
// only supplies what TRex's precompiled header would have declared before it: the unsigned_number concept, `ushort`, and inert glaze names for its
// JSON meta block.
#pragma once
#include <commons.pc.h>
#include <compare>
using ushort = unsigned short;
namespace cmn { template<typename T> concept unsigned_number = std::unsigned_integral<std::remove_cvref_t<T>>; }
#include_next <misc/bid.h>
