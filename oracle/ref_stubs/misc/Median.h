// Test infrastructure: misc/Median.h needs <queue> from TRex's precompiled header; the class itself is the REFERENCE'S OWN file, included from the checkout.
#pragma once
#include <commons.pc.h>
#include <queue>
#include_next <misc/Median.h>
