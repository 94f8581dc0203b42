// Test infrastructure: placeholder for tracker/python/PipelineRegistry.h; the inert registry is in python/Detection.h of this directory.
#pragma once
#include <python/Detection.h>
