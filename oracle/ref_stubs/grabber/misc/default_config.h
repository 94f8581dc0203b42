// Test infrastructure: placeholder for grabber/misc/default_config.h (FilterCache.cpp opens the namespace; meta_encoding_t itself is the reference's processing/encoding.h).
#pragma once
#include <commons.pc.h>
namespace grab::default_config {}
