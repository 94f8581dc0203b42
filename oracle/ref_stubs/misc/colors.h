// Test infrastructure: placeholder for a TRex header that Posture.cpp includes; what it needs from there is in tracking/Tracker.h of this directory.
#pragma once
#include <commons.pc.h>
