// Test infrastructure: placeholder for commons/common/misc/frame_t.h; the Frame_t look-alike (comparison, difference, the _f literal) is in commons.pc.h of this directory.
#pragma once
#include <commons.pc.h>
