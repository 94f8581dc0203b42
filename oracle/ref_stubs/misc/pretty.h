// Test infrastructure: placeholder for commons/common/misc/pretty.h.
#pragma once
#include <commons.pc.h>
