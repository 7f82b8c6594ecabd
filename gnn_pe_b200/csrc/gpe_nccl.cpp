// gpe_nccl.cpp -- NCCL bound at run time (dlopen), so that libgpe.so has no link-time dependency on it: a single-GPU
// user never touches NCCL, and inside a process that already carries an NCCL (PyTorch's bundled one) the same library
// is used instead of a second copy.  Only the handful of entry points the candidate exchange needs.
#include "gpe_nccl.h"

#include <dlfcn.h>

#include <mutex>

namespace gpe {

NcclApi &nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        void *h = nullptr;
        for (const char *n : names)
            if ((h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
        if (!h) {
            api.err = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "not found");
            return;
        }
#define GPE_SYM(field, name)                                                              \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name));                    \
    if (!api.field) { api.err = std::string("libnccl: missing symbol ") + name; return; }
        GPE_SYM(GetUniqueId, "ncclGetUniqueId")
        GPE_SYM(CommInitRank, "ncclCommInitRank")
        GPE_SYM(CommInitAll, "ncclCommInitAll")
        GPE_SYM(CommDestroy, "ncclCommDestroy")
        GPE_SYM(AllGather, "ncclAllGather")
        GPE_SYM(AllReduce, "ncclAllReduce")
        GPE_SYM(GroupStart, "ncclGroupStart")
        GPE_SYM(GroupEnd, "ncclGroupEnd")
        GPE_SYM(GetErrorString, "ncclGetErrorString")
        GPE_SYM(GetVersion, "ncclGetVersion")
#undef GPE_SYM
        api.ok = true;
    });
    return api;
}

}  // namespace gpe
