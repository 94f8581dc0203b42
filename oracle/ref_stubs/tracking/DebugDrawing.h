// Test infrastructure: placeholder for tracker/tracking/DebugDrawing.h (quote-included by Outline.cpp; nothing of it is needed for the functions under test).
#pragma once
#include <commons.pc.h>
