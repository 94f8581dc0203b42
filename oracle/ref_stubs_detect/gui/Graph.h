// Test infrastructure: placeholder for a TRex header that RawProcessing.cpp includes and uses nothing of on the path under test.
#pragma once
#include <commons.pc.h>
