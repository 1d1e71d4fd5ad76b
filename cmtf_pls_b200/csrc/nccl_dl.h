// NCCL bound at run time (dlopen) so that the library has no link-time dependency:
// inside a PyTorch process libnccl.so.2 is already mapped and is simply reused.
#pragma once

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>

namespace tpls {

struct NcclUniqueId {
    char internal[128];
};
typedef struct ncclComm* ncclComm_t;
constexpr int kNcclFloat64 = 8;  // ncclDataType_t::ncclFloat64
constexpr int kNcclSum = 0;      // ncclRedOp_t::ncclSum

struct NcclApi {
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    void* lib = nullptr;

    bool load(const char** why) {
        if (lib) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) {
            *why = "libnccl.so.2 not found (import torch first, or add it to LD_LIBRARY_PATH)";
            return false;
        }
        GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce || !GetErrorString) {
            *why = "libnccl is missing an expected symbol";
            return false;
        }
        return true;
    }
};

}  // namespace tpls
