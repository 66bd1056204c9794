// NCCL bound at run time (dlopen), for the in-process multi-GPU path of the host-pointer entry points (api.cu).
//
// libdcb200.so is loaded into two kinds of processes: the `clustering` binary / a C++ caller (nothing else brings NCCL),
// and Python next to torch, which ships its own libnccl.so.2.  Linking -lnccl would bind whichever copy the loader
// finds first and could shadow torch's; dlopen("libnccl.so.2") returns the copy already mapped when there is one and
// the system library otherwise.  Only types come from <nccl.h>; no NCCL symbol is referenced at link time.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>
#include <string>

namespace dcb {

struct NcclApi {
  void* handle = nullptr;
  bool ok = false;
  std::string why;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;

  template <class F>
  bool bind(F& fn, const char* name) {
    fn = reinterpret_cast<F>(dlsym(handle, name));
    if (!fn) why = std::string("libnccl lacks ") + name;
    return fn != nullptr;
  }
  void load() {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle) {
      const char* e = dlerror();
      why = std::string("dlopen(libnccl.so.2) failed: ") + (e ? e : "?");
      return;
    }
    ok = bind(GetErrorString, "ncclGetErrorString") && bind(CommInitAll, "ncclCommInitAll") && bind(CommDestroy, "ncclCommDestroy") &&
         bind(GroupStart, "ncclGroupStart") && bind(GroupEnd, "ncclGroupEnd") && bind(Broadcast, "ncclBroadcast") &&
         bind(AllGather, "ncclAllGather") && bind(GetVersion, "ncclGetVersion");
  }
};

inline NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] { api.load(); });
  return api;
}

}  // namespace dcb
