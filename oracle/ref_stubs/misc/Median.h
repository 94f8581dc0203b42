// Test infrastructure: CircularGraph.h includes misc/Median.h but uses nothing of it (the member is commented out).
#pragma once
#include <commons.pc.h>
