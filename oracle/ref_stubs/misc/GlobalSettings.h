// Test infrastructure: placeholder for commons/common/misc/GlobalSettings.h; the three settings Background.cpp caches and the callback registration are in misc/detail_bits.h.
#pragma once
#include <commons.pc.h>
