// Test infrastructure: placeholder for commons/common/misc/bid.h (the declarations the compiled files need are in processing/Background.h and processing/PVBlob.h of this directory).
#pragma once
#include <processing/Background.h>
