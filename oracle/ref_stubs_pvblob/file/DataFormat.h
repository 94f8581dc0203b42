// Test infrastructure: placeholder for commons/common/file/DataFormat.h -- processing/PVBlob.h only names two of its classes as friends.
#pragma once
#include <commons.pc.h>
namespace cmn { class DataFormat; class DataPackage; }
