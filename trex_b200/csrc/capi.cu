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

// Page-locked host frame buffers (the pool buffers::TileBuffers hands BackgroundSubtraction::apply, T/core/TileBuffers.h:13-22)
extern "C" int tb_host_alloc(size_t bytes, void **out)
{
    TB_REQUIRE(out && bytes > 0, TB_ERR_INVALID, "tb_host_alloc: null / empty argument");
    *out = nullptr;
    TB_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
    return TB_OK;
}
extern "C" int tb_host_free(void *p)
{
    if (p) TB_CUDA(cudaFreeHost(p));
    return TB_OK;
}
extern "C" int tb_host_register(void *p, size_t bytes)
{
    TB_REQUIRE(p && bytes > 0, TB_ERR_INVALID, "tb_host_register: null / empty argument");
    TB_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return TB_OK;
}
extern "C" int tb_host_unregister(void *p)
{
    TB_REQUIRE(p, TB_ERR_INVALID, "tb_host_unregister: null argument");
    TB_CUDA(cudaHostUnregister(p));
    return TB_OK;
}
