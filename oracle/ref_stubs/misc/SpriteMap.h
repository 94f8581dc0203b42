// Test infrastructure: placeholder for commons/common/misc/SpriteMap.h (AveragingAccumulator.cpp includes it and uses nothing of it).
#pragma once
#include <commons.pc.h>
