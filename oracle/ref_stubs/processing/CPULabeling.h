// Test infrastructure: placeholder for commons/common/processing/CPULabeling.h; the declarations PixelTree.cpp needs are in processing/pixeltree_standins.h.
#pragma once
#include <processing/pixeltree_standins.h>
