// Test infrastructure: inert stand-ins for TRex's Timing / TakeTiming profiling helpers (commons/common/misc/Timer.h).
#pragma once
#include <commons.pc.h>
namespace cmn {
struct Timing { Timing(const char *, double = 0) {} };
struct TakeTiming { explicit TakeTiming(Timing&) {} };
}
struct Timer { double elapsed() const { return 0; } void reset() {} };      // FilterCache.cpp's message throttle
namespace cmn {
}
