// Multi-GPU plumbing: NCCL (resolved at run time with dlopen so the library neither links a second NCCL
// next to torch's nor needs one on a single-GPU box), contiguous row-block ownership, halo exchange of
// interface node layers and all-reduced CG scalars.  New relative to the reference, which has no
// distributed code at all.
#pragma once
#include <dlfcn.h>

#include <string>
#include <vector>

#include "ikb_internal.cuh"

namespace ikb {

struct NcclId {
  char internal[128];
};
typedef int (*nccl_get_unique_id_t)(NcclId*);
typedef int (*nccl_comm_init_rank_t)(void**, int, NcclId, int);
typedef int (*nccl_comm_destroy_t)(void*);
typedef int (*nccl_all_reduce_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_all_gather_t)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*nccl_send_t)(const void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_recv_t)(void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_group_t)(void);
typedef const char* (*nccl_err_t)(int);

constexpr int NCCL_INT64 = 4, NCCL_FLOAT64 = 8, NCCL_SUM = 0;

struct NcclApi {
  void* lib = nullptr;
  nccl_get_unique_id_t getUniqueId = nullptr;
  nccl_comm_init_rank_t commInitRank = nullptr;
  nccl_comm_destroy_t commDestroy = nullptr;
  nccl_all_reduce_t allReduce = nullptr;
  nccl_all_gather_t allGather = nullptr;
  nccl_send_t send = nullptr;
  nccl_recv_t recv = nullptr;
  nccl_group_t groupStart = nullptr, groupEnd = nullptr;
  nccl_err_t errorString = nullptr;

  bool load(std::string& err) {
    if (lib) return true;
    // prefer a copy the process already holds (torch's bundled NCCL), else the system one
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD);
      if (lib) break;
    }
    if (!lib)
      for (const char* n : names) {
        lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (lib) break;
      }
    if (!lib) {
      err = std::string("cannot load libnccl.so.2: ") + dlerror();
      return false;
    }
#define IKB_SYM(field, name)                                  \
  field = reinterpret_cast<decltype(field)>(dlsym(lib, name)); \
  if (!field) {                                               \
    err = std::string("missing NCCL symbol ") + name;         \
    lib = nullptr;                                            \
    return false;                                             \
  }
    IKB_SYM(getUniqueId, "ncclGetUniqueId")
    IKB_SYM(commInitRank, "ncclCommInitRank")
    IKB_SYM(commDestroy, "ncclCommDestroy")
    IKB_SYM(allReduce, "ncclAllReduce")
    IKB_SYM(allGather, "ncclAllGather")
    IKB_SYM(send, "ncclSend")
    IKB_SYM(recv, "ncclRecv")
    IKB_SYM(groupStart, "ncclGroupStart")
    IKB_SYM(groupEnd, "ncclGroupEnd")
    IKB_SYM(errorString, "ncclGetErrorString")
#undef IKB_SYM
    return true;
  }
};

inline NcclApi& nccl() {
  static NcclApi api;
  return api;
}

// Intervals (in nodes) this rank sends to / receives from a peer: my owned rows that the peer's columns
// touch, and the peer's owned rows that my columns touch.  Pure host logic (tested on CPU).
inline void haloIntervals(int64_t ownB, int64_t ownE, int64_t needB, int64_t needE, int64_t peerOwnB, int64_t peerOwnE,
                          int64_t peerNeedB, int64_t peerNeedE, int64_t& sendB, int64_t& sendE, int64_t& recvB,
                          int64_t& recvE) {
  sendB = ownB > peerNeedB ? ownB : peerNeedB;
  sendE = ownE < peerNeedE ? ownE : peerNeedE;
  recvB = peerOwnB > needB ? peerOwnB : needB;
  recvE = peerOwnE < needE ? peerOwnE : needE;
  if (sendE < sendB) sendE = sendB;
  if (recvE < recvB) recvE = recvB;
  (void)needE;
}

}  // namespace ikb
