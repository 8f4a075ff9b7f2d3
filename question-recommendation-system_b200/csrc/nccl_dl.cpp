// csrc/nccl_dl.cpp -- see nccl_dl.hpp.
#include "nccl_dl.hpp"

#include <dlfcn.h>

#include <mutex>
#include <string>

#include "engine.hpp"

namespace mfb200 {

const NcclApi *nccl_api() {
    static NcclApi api;
    static bool ok = false, tried = false;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (tried) return ok ? &api : nullptr;
    tried = true;
    void *h = nullptr;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        set_error(std::string("cannot load NCCL (needed for more than one GPU): ") + dlerror());
        return nullptr;
    }
#define MFB_SYM(field, name)                                        \
    *(void **)(&api.field) = dlsym(h, name);                        \
    if (!api.field) {                                               \
        set_error(std::string("NCCL symbol missing: ") + name);     \
        return nullptr;                                             \
    }
    MFB_SYM(GetUniqueId, "ncclGetUniqueId")
    MFB_SYM(CommInitRank, "ncclCommInitRank")
    MFB_SYM(CommDestroy, "ncclCommDestroy")
    MFB_SYM(GetErrorString, "ncclGetErrorString")
    MFB_SYM(GroupStart, "ncclGroupStart")
    MFB_SYM(GroupEnd, "ncclGroupEnd")
    MFB_SYM(Send, "ncclSend")
    MFB_SYM(Recv, "ncclRecv")
    MFB_SYM(AllReduce, "ncclAllReduce")
    MFB_SYM(AllGather, "ncclAllGather")
#undef MFB_SYM
    ok = true;
    return &api;
}

}  // namespace mfb200
