// nccl_dyn.h — the one collective of the path, inside the library: a run-time binding of NCCL (dlopen, no link- or
// build-time dependency) for the all-reduce of [cost, grad…] over the ranks of a one-process-per-GPU job
// (BASELINE north_star: "a single NCCL allreduce over NVLink of the upper-level loss and gradient per trust-region
// step"; the reference sums the per-image terms in `for i = 1:O`, /root/reference/src/TVLearningFunctionVec.jl:72-83).
// Only the five entry points used are declared; their signatures are NCCL's stable C ABI (nccl.h, 2.x):
// ncclUniqueId is a 128-byte struct passed by value, ncclDouble = 8 (ncclFloat64), ncclSum = 0.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdlib>
#include <mutex>
#include <string>

namespace bpltv {

struct NcclId { char internal[128]; };

struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int *) = nullptr;
    std::string err;
    std::mutex mu;

    // BPLTV_NCCL_LIB names the library; otherwise the soname (a process that already loaded NCCL — e.g. through
    // torch — gets that copy), then the unversioned name
    bool load()
    {
        std::lock_guard<std::mutex> lock(mu);      // ranks of one process (host threads) may arrive together
        if (handle) return true;
        const char *names[3] = {std::getenv("BPLTV_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (!n || !*n) continue;
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
            err = dlerror();
        }
        if (!handle) return false;
        auto sym = [&](const char *s) -> void * {
            void *p = dlsym(handle, s);
            if (!p) err = std::string("missing symbol ") + s;
            return p;
        };
        GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(sym("ncclGetUniqueId"));
        CommInitRank = reinterpret_cast<decltype(CommInitRank)>(sym("ncclCommInitRank"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
        AllReduce = reinterpret_cast<decltype(AllReduce)>(sym("ncclAllReduce"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
        GetVersion = reinterpret_cast<decltype(GetVersion)>(sym("ncclGetVersion"));
        if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce || !GetErrorString) {
            dlclose(handle);
            handle = nullptr;
            return false;
        }
        return true;
    }
    const char *what(int rc) const { return GetErrorString ? GetErrorString(rc) : "NCCL error"; }
};

static inline NcclApi &nccl_api()
{
    static NcclApi api;
    return api;
}

}  // namespace bpltv
