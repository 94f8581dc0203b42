// Test infrastructure: placeholder for commons/common/processing/LuminanceGrid.h (PixelTree.cpp includes it; the grid variant of its templates is commented out upstream).
#pragma once
#include <commons.pc.h>
