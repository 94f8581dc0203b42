// Error plumbing and library-level entry points of the C ABI (include/trexb200.h).
#include "common.h"

namespace tb {
static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
const char *get_error() { return g_err.c_str(); }
}  // namespace tb

extern "C" const char *tb_last_error(void) { return tb::get_error(); }
extern "C" int tb_abi_version(void) { return TB_ABI_VERSION; }
extern "C" int tb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
