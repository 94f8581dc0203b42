// Test infrastructure: placeholder for commons/common/misc/Grid.h -- processing/PVBlob.h includes it and uses nothing of it.
#pragma once
#include <commons.pc.h>
