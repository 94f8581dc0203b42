// Test infrastructure: placeholder for commons/common/misc/bid.h; the declarations PixelTree.cpp needs are in processing/pixeltree_standins.h.
#pragma once
#include <processing/pixeltree_standins.h>
