// NCCL is loaded lazily with dlopen so that single-GPU use needs no libnccl and so that the
// library shares whatever libnccl.so.2 the host process (e.g. torch) has already mapped.
#include <dlfcn.h>
#include <stddef.h>
#include <string.h>

#include "nccl_dl.h"

static NcclApi g_api;
static bool g_tried = false;

const NcclApi* nccl_api() {
    if (g_tried) return g_api.ok ? &g_api : nullptr;
    g_tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    void* h = nullptr;
    for (int i = 0; names[i] && !h; ++i) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!h) return nullptr;
    g_api.GetUniqueId = (int (*)(NcclUid*))dlsym(h, "ncclGetUniqueId");
    g_api.CommInitRank = (int (*)(void**, int, NcclUid, int))dlsym(h, "ncclCommInitRank");
    g_api.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
    g_api.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, void*))dlsym(h, "ncclAllReduce");
    g_api.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    g_api.ok = g_api.GetUniqueId && g_api.CommInitRank && g_api.CommDestroy && g_api.AllReduce;
    return g_api.ok ? &g_api : nullptr;
}
