// Symmetric heap runtime: physical allocations (CUDA VMM) that every rank of a node-local process group maps into its own
// address space for in-kernel P2P loads/stores over NVLink 5, plus NVSwitch multicast (NVLS) objects for multimem.* ops.
//
// Replaces, in one place, what the reference spreads over three extensions: the ncclMemAlloc pluggable allocator
// (apex/contrib/csrc/nccl_allocator/NCCLAllocator.cpp:17-38), peer_memory's cudaIpc pool
// (apex/contrib/csrc/peer_memory/peer_memory_cuda.cu:318-354) and groupbn's IPC buffers (apex/contrib/csrc/groupbn/ipc.cu:53-114).
//
// Handle exchange (POSIX file descriptors) and ordering are done by python (apex_b200/parallel/symmetric.py) over the
// torch.distributed bootstrap; this file only talks to the driver. libcuda is resolved at run time through
// cudaGetDriverEntryPoint, so the library still loads on a machine without a driver.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#define AB_API extern "C" __attribute__((visibility("default")))

namespace {

template <typename F>
F drv(const char* name) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult st;
  if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess) return nullptr;
  return reinterpret_cast<F>(fn);
}

#define DRV(name) static auto p_##name = drv<decltype(&name)>(#name); if (!p_##name) return -1000
#define CK(call)                                              \
  do {                                                        \
    CUresult r__ = (call);                                    \
    if (r__ != CUDA_SUCCESS) { last_err = (int)r__; return (int)r__; } \
  } while (0)

thread_local int last_err = 0;

CUmemAllocationProp local_prop(int device) {
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}

int map_rw(CUmemGenericAllocationHandle h, size_t bytes, int device, size_t gran, CUdeviceptr* out) {
  DRV(cuMemAddressReserve); DRV(cuMemMap); DRV(cuMemSetAccess);
  CUdeviceptr va = 0;
  CK(p_cuMemAddressReserve(&va, bytes, gran, 0, 0));
  CK(p_cuMemMap(va, bytes, 0, h, 0));
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  CK(p_cuMemSetAccess(va, bytes, &acc, 1));
  *out = va;
  return 0;
}

}  // namespace

AB_API int ab_symm_last_error() { return last_err; }

// Round `bytes` up to what the driver wants for shareable physical allocations (and multicast binding when mc != 0).
AB_API int ab_symm_granularity(int device, int num_devices_for_mc, uint64_t* gran_out) {
  DRV(cuMemGetAllocationGranularity);
  CUmemAllocationProp prop = local_prop(device);
  size_t g = 0;
  CK(p_cuMemGetAllocationGranularity(&g, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  if (num_devices_for_mc > 1) {
    static auto p_mcg = drv<decltype(&cuMulticastGetGranularity)>("cuMulticastGetGranularity");
    if (p_mcg) {
      CUmulticastObjectProp mp;
      memset(&mp, 0, sizeof(mp));
      mp.numDevices = num_devices_for_mc;
      mp.size = g;
      mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
      size_t mg = 0;
      if (p_mcg(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mg > g) g = mg;
    }
  }
  *gran_out = g;
  return 0;
}

// Allocate `bytes` (already a multiple of the granularity) of device memory, map it locally, export a POSIX fd.
AB_API int ab_symm_alloc(int device, uint64_t bytes, uint64_t gran, uint64_t* handle_out, uint64_t* ptr_out, int* fd_out) {
  DRV(cuMemCreate); DRV(cuMemExportToShareableHandle);
  cudaSetDevice(device);
  cudaFree(0);  // make sure the primary context exists
  CUmemAllocationProp prop = local_prop(device);
  CUmemGenericAllocationHandle h;
  CK(p_cuMemCreate(&h, bytes, &prop, 0));
  CUdeviceptr va = 0;
  int rc = map_rw(h, bytes, device, gran, &va);
  if (rc) return rc;
  int fd = -1;
  CK(p_cuMemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  *handle_out = (uint64_t)h;
  *ptr_out = (uint64_t)va;
  *fd_out = fd;
  return 0;
}

// Map a peer's allocation (fd received from that peer) read/write into this process for `device`.
AB_API int ab_symm_import(int device, int fd, uint64_t bytes, uint64_t gran, uint64_t* handle_out, uint64_t* ptr_out) {
  DRV(cuMemImportFromShareableHandle);
  cudaSetDevice(device);
  CUmemGenericAllocationHandle h;
  CK(p_cuMemImportFromShareableHandle(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  CUdeviceptr va = 0;
  int rc = map_rw(h, bytes, device, gran, &va);
  if (rc) return rc;
  *handle_out = (uint64_t)h;
  *ptr_out = (uint64_t)va;
  return 0;
}

AB_API int ab_symm_free(uint64_t handle, uint64_t ptr, uint64_t bytes) {
  DRV(cuMemUnmap); DRV(cuMemAddressFree); DRV(cuMemRelease);
  if (ptr) { p_cuMemUnmap((CUdeviceptr)ptr, bytes); p_cuMemAddressFree((CUdeviceptr)ptr, bytes); }
  if (handle) p_cuMemRelease((CUmemGenericAllocationHandle)handle);
  return 0;
}

// ---- NVSwitch multicast (NVLS) ----------------------------------------------------------------------------------------
AB_API int ab_mc_supported(int device) {
  int v = 0;
  DRV(cuDeviceGetAttribute);
  if (p_cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, device) != CUDA_SUCCESS) return 0;
  return v;
}

// Rank 0 of the group: create the multicast object and export it.
AB_API int ab_mc_create(int num_devices, uint64_t bytes, uint64_t* handle_out, int* fd_out) {
  DRV(cuMulticastCreate); DRV(cuMemExportToShareableHandle);
  CUmulticastObjectProp mp;
  memset(&mp, 0, sizeof(mp));
  mp.numDevices = num_devices;
  mp.size = bytes;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  CUmemGenericAllocationHandle h;
  CK(p_cuMulticastCreate(&h, &mp));
  int fd = -1;
  CK(p_cuMemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  *handle_out = (uint64_t)h;
  *fd_out = fd;
  return 0;
}

AB_API int ab_mc_import(int fd, uint64_t* handle_out) {
  DRV(cuMemImportFromShareableHandle);
  CUmemGenericAllocationHandle h;
  CK(p_cuMemImportFromShareableHandle(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  *handle_out = (uint64_t)h;
  return 0;
}

AB_API int ab_mc_add_device(uint64_t mc_handle, int device) {
  DRV(cuMulticastAddDevice);
  CK(p_cuMulticastAddDevice((CUmemGenericAllocationHandle)mc_handle, device));
  return 0;
}

// After EVERY rank has added its device: bind this rank's physical allocation at offset 0 and map the multicast VA.
AB_API int ab_mc_bind_map(uint64_t mc_handle, uint64_t mem_handle, int device, uint64_t bytes, uint64_t gran, uint64_t* mc_ptr_out) {
  DRV(cuMulticastBindMem);
  CK(p_cuMulticastBindMem((CUmemGenericAllocationHandle)mc_handle, 0, (CUmemGenericAllocationHandle)mem_handle, 0, bytes, 0));
  CUdeviceptr va = 0;
  int rc = map_rw((CUmemGenericAllocationHandle)mc_handle, bytes, device, gran, &va);
  if (rc) return rc;
  *mc_ptr_out = (uint64_t)va;
  return 0;
}

// ---- legacy CUDA-IPC fallback (no multicast) ---------------------------------------------------------------------------
AB_API int ab_ipc_alloc(int device, uint64_t bytes, uint64_t* ptr_out, void* handle64_out) {
  cudaSetDevice(device);
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return (int)e;
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) return (int)e;
  memcpy(handle64_out, &h, sizeof(h));
  *ptr_out = (uint64_t)p;
  return 0;
}

AB_API int ab_ipc_open(int device, const void* handle64, uint64_t* ptr_out) {
  cudaSetDevice(device);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return (int)e;
  *ptr_out = (uint64_t)p;
  return 0;
}

AB_API int ab_ipc_close(uint64_t ptr) { return (int)cudaIpcCloseMemHandle((void*)ptr); }
AB_API int ab_ipc_free(uint64_t ptr) { return (int)cudaFree((void*)ptr); }

AB_API int ab_close_fd(int fd) { return close(fd); }

// ---- pluggable allocator over the shareable heap -----------------------------------------------------------------------
// torch.cuda.memory.CUDAPluggableAllocator(<this library>, "ab_symm_pool_malloc", "ab_symm_pool_free") -> torch.cuda.MemPool: every
// tensor allocated inside the pool lives in a shareable VMM allocation, so it can be peer-mapped (and multicast-bound) AFTER the fact
// with ab_symm_pool_export + ab_symm_import on the peers. The reference's counterpart wraps ncclMemAlloc / ncclMemFree
// (apex/contrib/csrc/nccl_allocator/NCCLAllocator.cpp:17-38) so that NCCL can register user buffers; here the consumers are this
// library's own in-kernel collectives.
#include <map>
#include <mutex>

namespace {
struct PoolBlock { CUmemGenericAllocationHandle h; size_t bytes; int device; };
std::map<uintptr_t, PoolBlock> g_pool;   // base address -> block
std::mutex g_pool_mu;
}  // namespace

extern "C" __attribute__((visibility("default"))) void* ab_symm_pool_malloc(ssize_t size, int device, cudaStream_t) {
  if (size <= 0) return nullptr;
  uint64_t gran = 0;
  if (ab_symm_granularity(device, 1, &gran) != 0 || gran == 0) return nullptr;
  const uint64_t bytes = ((uint64_t)size + gran - 1) / gran * gran;
  uint64_t h = 0, p = 0;
  int fd = -1;
  if (ab_symm_alloc(device, bytes, gran, &h, &p, &fd) != 0) return nullptr;
  if (fd >= 0) close(fd);   // re-exported on demand (ab_symm_pool_export)
  std::lock_guard<std::mutex> lk(g_pool_mu);
  g_pool[(uintptr_t)p] = PoolBlock{(CUmemGenericAllocationHandle)h, (size_t)bytes, device};
  return (void*)(uintptr_t)p;
}

extern "C" __attribute__((visibility("default"))) void ab_symm_pool_free(void* ptr, ssize_t, int, cudaStream_t) {
  PoolBlock b;
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    auto it = g_pool.find((uintptr_t)ptr);
    if (it == g_pool.end()) return;
    b = it->second;
    g_pool.erase(it);
  }
  ab_symm_free((uint64_t)b.h, (uint64_t)(uintptr_t)ptr, (uint64_t)b.bytes);
}

// Block that contains `ptr`: base address, size, and a fresh POSIX fd of its physical allocation (the caller closes it after its peers
// have imported it). Returns -3 when the pointer does not belong to the pool.
AB_API int ab_symm_pool_export(const void* ptr, uint64_t* base_out, uint64_t* bytes_out, int* fd_out) {
  DRV(cuMemExportToShareableHandle);
  PoolBlock b;
  uintptr_t base = 0;
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    auto it = g_pool.upper_bound((uintptr_t)ptr);
    if (it == g_pool.begin()) return -3;
    --it;
    if ((uintptr_t)ptr >= it->first + it->second.bytes) return -3;
    base = it->first;
    b = it->second;
  }
  int fd = -1;
  CK(p_cuMemExportToShareableHandle(&fd, b.h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  *base_out = (uint64_t)base; *bytes_out = (uint64_t)b.bytes; *fd_out = fd;
  return 0;
}

AB_API int ab_symm_pool_blocks() {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  return (int)g_pool.size();
}
