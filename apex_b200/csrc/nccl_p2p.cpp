// A communicator of our own for neighbour exchanges (reference apex/contrib/csrc/nccl_p2p/nccl_p2p_cuda.cu:34-128: ncclGetUniqueId /
// ncclCommInitRank / grouped ncclSend + ncclRecv on the current stream). NCCL is resolved at run time from the library torch already
// loaded (no link-time dependency, no headers): only the handful of entry points below are used.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>
#include <cstring>
#include <mutex>
#include <vector>

#define AB_API extern "C" __attribute__((visibility("default")))

namespace {

struct UniqueId { char internal[128]; };
typedef void* Comm;
typedef int (*GetUniqueIdFn)(UniqueId*);
typedef int (*CommInitRankFn)(Comm*, int, UniqueId, int);
typedef int (*CommDestroyFn)(Comm);
typedef int (*GroupFn)();
typedef int (*SendFn)(const void*, size_t, int, int, Comm, cudaStream_t);
typedef int (*RecvFn)(void*, size_t, int, int, Comm, cudaStream_t);
typedef const char* (*ErrStrFn)(int);

struct Api {
  void* lib = nullptr;
  GetUniqueIdFn get_id = nullptr; CommInitRankFn init = nullptr; CommDestroyFn destroy = nullptr;
  GroupFn gstart = nullptr, gend = nullptr; SendFn send = nullptr; RecvFn recv = nullptr;
  bool ok() const { return lib && get_id && init && destroy && gstart && gend && send && recv; }
};
Api g_api;
std::mutex g_mu;
std::vector<Comm> g_comms;

int load(const char* path) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_api.ok()) return 0;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy torch.distributed already mapped, if any
  if (!h && path && *path) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return -1001;
  g_api.lib = h;
  g_api.get_id = (GetUniqueIdFn)dlsym(h, "ncclGetUniqueId");
  g_api.init = (CommInitRankFn)dlsym(h, "ncclCommInitRank");
  g_api.destroy = (CommDestroyFn)dlsym(h, "ncclCommDestroy");
  g_api.gstart = (GroupFn)dlsym(h, "ncclGroupStart");
  g_api.gend = (GroupFn)dlsym(h, "ncclGroupEnd");
  g_api.send = (SendFn)dlsym(h, "ncclSend");
  g_api.recv = (RecvFn)dlsym(h, "ncclRecv");
  return g_api.ok() ? 0 : -1002;
}

}  // namespace

AB_API int ab_nccl_load(const char* path) { return load(path); }

// out: 128 bytes
AB_API int ab_nccl_unique_id(void* out) {
  if (!g_api.ok()) return -1001;
  UniqueId id;
  const int rc = g_api.get_id(&id);
  if (rc) return 3000 + rc;
  memcpy(out, id.internal, 128);
  return 0;
}

// -> *handle = index of the new communicator (collective over the `nranks` callers sharing `id`)
AB_API int ab_nccl_comm_init(const void* id128, int rank, int nranks, int* handle) {
  if (!g_api.ok()) return -1001;
  UniqueId id;
  memcpy(id.internal, id128, 128);
  Comm c = nullptr;
  const int rc = g_api.init(&c, nranks, id, rank);
  if (rc) return 3000 + rc;
  std::lock_guard<std::mutex> lk(g_mu);
  g_comms.push_back(c);
  *handle = (int)g_comms.size() - 1;
  return 0;
}

AB_API int ab_nccl_comm_destroy(int handle) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (handle < 0 || handle >= (int)g_comms.size() || !g_comms[handle]) return -3;
  g_api.destroy(g_comms[handle]);
  g_comms[handle] = nullptr;
  return 0;
}

// One grouped exchange on `st`: send `nbytes` from send_lo to peer_lo and from send_hi to peer_hi, receive as many into recv_lo / recv_hi.
// A negative peer skips that side (ends of an open chain).
AB_API int ab_nccl_exchange(int handle, int peer_lo, int peer_hi, const void* send_lo, void* recv_lo, const void* send_hi, void* recv_hi,
                            long long nbytes, cudaStream_t st) {
  Comm c;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (handle < 0 || handle >= (int)g_comms.size() || !g_comms[handle]) return -3;
    c = g_comms[handle];
  }
  int rc = g_api.gstart();
  if (rc) return 3000 + rc;
  const int kInt8 = 0;
  if (peer_lo >= 0) {
    if ((rc = g_api.send(send_lo, (size_t)nbytes, kInt8, peer_lo, c, st))) { g_api.gend(); return 3000 + rc; }
    if ((rc = g_api.recv(recv_lo, (size_t)nbytes, kInt8, peer_lo, c, st))) { g_api.gend(); return 3000 + rc; }
  }
  if (peer_hi >= 0) {
    if ((rc = g_api.send(send_hi, (size_t)nbytes, kInt8, peer_hi, c, st))) { g_api.gend(); return 3000 + rc; }
    if ((rc = g_api.recv(recv_hi, (size_t)nbytes, kInt8, peer_hi, c, st))) { g_api.gend(); return 3000 + rc; }
  }
  rc = g_api.gend();
  return rc ? 3000 + rc : 0;
}
