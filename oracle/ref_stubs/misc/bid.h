// Test infrastructure: placeholder for commons/common/misc/bid.h (PixelTree.h includes it; blob ids are not used by the functions under test).
#pragma once
#include <commons.pc.h>
