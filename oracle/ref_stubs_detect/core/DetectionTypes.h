// Test infrastructure: placeholder for tracker/core/DetectionTypes.h; the ObjectDetectionType look-alike is in python/Detection.h of this directory.
#pragma once
#include <python/Detection.h>
