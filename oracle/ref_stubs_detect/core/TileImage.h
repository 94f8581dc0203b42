// Test infrastructure: placeholder for tracker/core/TileImage.h; the TileImage / SegmentationData look-alikes are in python/Detection.h of this directory.
#pragma once
#include <python/Detection.h>
