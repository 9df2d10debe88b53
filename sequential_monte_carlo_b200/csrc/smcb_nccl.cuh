// NCCL bound at run time (dlopen), so that libsmcb200.so loads on machines without NCCL and a
// single-GPU user never touches it.  Only the θ-sharded samplers (smcb_comm_init) need it:
// one all-gather of M/G log-likelihoods per step and grouped send/recv of whole state clouds after a
// θ-resample, both on the context's stream, straight from / to device buffers.
//
// Library resolution: $SMCB_NCCL_LIB, else "libnccl.so.2" by soname — inside a process that already
// loaded NCCL (e.g. PyTorch's bundled copy) the loader hands back that same copy.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>
#include <stdint.h>

#include <cstdlib>
#include <string>

#include "smcb_common.cuh"

namespace smcb {

// the slice of nccl.h this library uses (ABI-stable since NCCL 2.0)
struct NcclUniqueId {
  char internal[128];
};
typedef struct ncclComm* NcclComm;
enum : int { kNcclSuccess = 0 };
enum : int { kNcclUint8 = 1, kNcclInt32 = 2, kNcclFloat64 = 8 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetVersion)(int*) = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
};

inline const NcclApi& nccl_api() {
  static NcclApi api;
  static bool tried = false;
  static std::string why;
  if (!tried) {
    tried = true;
    const char* env = std::getenv("SMCB_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
      why = dlerror();
    }
    if (api.handle) {
      auto sym = [&](const char* s) -> void* {
        void* p = dlsym(api.handle, s);
        if (!p) why = std::string("missing symbol ") + s;
        return p;
      };
      api.GetVersion = reinterpret_cast<int (*)(int*)>(sym("ncclGetVersion"));
      api.GetUniqueId = reinterpret_cast<int (*)(NcclUniqueId*)>(sym("ncclGetUniqueId"));
      api.CommInitRank = reinterpret_cast<int (*)(NcclComm*, int, NcclUniqueId, int)>(sym("ncclCommInitRank"));
      api.CommDestroy = reinterpret_cast<int (*)(NcclComm)>(sym("ncclCommDestroy"));
      api.GetErrorString = reinterpret_cast<const char* (*)(int)>(sym("ncclGetErrorString"));
      api.AllGather = reinterpret_cast<int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t)>(sym("ncclAllGather"));
      api.Send = reinterpret_cast<int (*)(const void*, size_t, int, int, NcclComm, cudaStream_t)>(sym("ncclSend"));
      api.Recv = reinterpret_cast<int (*)(void*, size_t, int, int, NcclComm, cudaStream_t)>(sym("ncclRecv"));
      api.GroupStart = reinterpret_cast<int (*)()>(sym("ncclGroupStart"));
      api.GroupEnd = reinterpret_cast<int (*)()>(sym("ncclGroupEnd"));
      if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.GetErrorString || !api.AllGather || !api.Send ||
          !api.Recv || !api.GroupStart || !api.GroupEnd) {
        dlclose(api.handle);
        api.handle = nullptr;
      }
    }
  }
  if (!api.handle) throw Error{SMCB_ERR_NCCL, "NCCL is not available (set SMCB_NCCL_LIB to libnccl.so.2): " + why};
  return api;
}

#define SMCB_NCCL_TRY(expr)                                                                       \
  do {                                                                                            \
    int _r = (expr);                                                                              \
    if (_r != ::smcb::kNcclSuccess)                                                               \
      throw ::smcb::Error{::SMCB_ERR_NCCL, std::string(#expr) + ": " + ::smcb::nccl_api().GetErrorString(_r)}; \
  } while (0)

// one communicator per context: rank r of G owns θ-particles [r M/G, (r+1) M/G)
struct Comm {
  NcclComm comm = nullptr;
  int rank = 0, nranks = 1;
  bool active() const { return nranks > 1; }
};

}  // namespace smcb
