#pragma once
#include <stddef.h>

struct NcclUid { char internal[128]; };  // == ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
enum { kNcclSum = 0, kNcclDouble = 8 };  // ncclSum, ncclFloat64 (nccl.h 2.27)

struct NcclApi {
    bool ok = false;
    int (*GetUniqueId)(NcclUid*) = nullptr;
    int (*CommInitRank)(void**, int, NcclUid, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, void*) = nullptr;  // (.., dtype, op, comm, stream)
    const char* (*GetErrorString)(int) = nullptr;
};

const NcclApi* nccl_api();
