// Test infrastructure: placeholder for a TRex header that the compiled reference files include but need nothing from here.
#pragma once
#include <commons.pc.h>
