// nccl_comm.cuh - the data-parallel exchange of the GAN step behind the C ABI (fsr_nccl_* in include/fsr_b200.h).
//
// The reference has no collective at all ("Distributed comm backend: None"); the engine adds exactly two
// per step - sum of the flat fp32 discriminator gradient after trainer.py:180 and of the generator gradient after
// trainer.py:195.  They are issued from THIS library (ncclAllReduce on the caller's stream) so that they can sit
// inside the captured CUDA graph of the step, on a side stream that overlaps the VGG passes (engine.py).
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): when PyTorch is in the process its bundled NCCL is the one that is
// found (RTLD_NOLOAD first), so the process never holds two NCCL copies; without any libnccl the entry points return
// FSR_ERR_NO_NCCL and single-GPU use is unaffected.  Minimal declarations below mirror <nccl.h> (stable ABI since 2.x).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <mutex>

namespace fsr {

struct NcclUniqueId { char internal[128]; };   // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef void* NcclComm;                        // ncclComm_t
constexpr int kNcclFloat32 = 7;                // ncclFloat32
constexpr int kNcclSum = 0;                    // ncclSum

struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

inline const NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // PyTorch's bundled copy if already mapped
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(h, "ncclAllReduce"));
    api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(dlsym(h, "ncclBroadcast"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(dlsym(h, "ncclGetVersion"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.Broadcast && api.CommDestroy;
  });
  return api;
}

}  // namespace fsr
