// Test infrastructure: placeholder for commons/common/gui/DrawCVBase.h; it only has to bring the (real) gui/Transform.h into Outline.cpp.
#pragma once
#include <commons.pc.h>
#include <gui/Transform.h>
