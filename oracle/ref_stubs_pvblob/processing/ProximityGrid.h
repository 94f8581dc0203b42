// Test infrastructure: placeholder for commons/common/processing/ProximityGrid.h -- processing/PVBlob.h includes it and uses nothing of it.
#pragma once
#include <commons.pc.h>
