// NCCL binding of the multi-GPU witness path: the collectives live behind the C ABI, on the engine's own
// streams.  libnccl.so.2 is resolved at run time (dlopen) the first time a communicator is created, so the library loads -- and the
// single-GPU path works -- on hosts without NCCL; inside a process that already mapped an NCCL (e.g. PyTorch's bundled one) the
// same copy is reused.  Only types come from <nccl.h>; no NCCL symbol is linked.
#pragma once
#include <dlfcn.h>
#include <nccl.h>
#include <mutex>
#include <string>

namespace eagen {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string load_error;
    bool ok = false;

    static NcclApi& get() {
        static NcclApi api;
        static std::once_flag once;
        std::call_once(once, [] { api.load(); });
        return api;
    }

private:
    void load() {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { load_error = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "?"); return; }
        auto sym = [&](const char* name) -> void* {
            void* p = dlsym(h, name);
            if (!p && load_error.empty()) load_error = std::string("NCCL symbol missing: ") + name;
            return p;
        };
        GetUniqueId = (decltype(GetUniqueId))sym("ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))sym("ncclCommInitRank");
        CommInitAll = (decltype(CommInitAll))sym("ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
        AllGather = (decltype(AllGather))sym("ncclAllGather");
        GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
        GetVersion = (decltype(GetVersion))sym("ncclGetVersion");
        GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
        ok = load_error.empty();
    }
};

}  // namespace eagen
