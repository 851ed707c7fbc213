// Tensor <-> file transfer runtime (the storage side of checkpointing).
//
// Reference: apex/contrib/csrc/gpu_direct_storage/gds.cpp:45-165 — cuFileRead / cuFileWrite of a tensor's storage plus "no_gds"
// fallbacks that bounce through an unpinned CPU tensor. cuFile is not available in this image, so this is the bounce path done
// properly: two pinned staging buffers, the device copy of chunk i+1 overlapped with the file I/O of chunk i (writes), the file read
// of chunk i+1 overlapped with the host-to-device copy of chunk i (reads). Host tensors go straight through pread / pwrite.
// Torch-free: python passes raw pointers, a byte count, a file offset and the stream (apex_b200/contrib/gpu_direct_storage).
#include <cuda_runtime.h>
#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdlib.h>
#include <unistd.h>

#include <mutex>

#define AB_API extern "C" __attribute__((visibility("default")))

namespace {

constexpr size_t kChunk = 32u << 20;  // staging chunk: large enough to hide pwrite / pread latency, small enough to stay resident

struct Staging {
  void* buf[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  std::mutex mu;  // one transfer at a time per process: the two buffers ARE the pipeline
  int ensure() {
    for (int i = 0; i < 2; i++) {
      if (!buf[i]) {
        cudaError_t e = cudaHostAlloc(&buf[i], kChunk, cudaHostAllocDefault);
        if (e != cudaSuccess) { buf[i] = nullptr; return (int)e; }
      }
      if (!ev[i]) {
        cudaError_t e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
        if (e != cudaSuccess) { ev[i] = nullptr; return (int)e; }
      }
    }
    return 0;
  }
};
Staging g_stage;

// pwrite / pread until `n` bytes are done (short transfers and EINTR are normal on network and FUSE file systems)
long long write_all(int fd, const char* p, size_t n, off_t off) {
  size_t done = 0;
  while (done < n) {
    ssize_t r = pwrite(fd, p + done, n - done, off + (off_t)done);
    if (r < 0) { if (errno == EINTR) continue; return -(long long)errno; }
    done += (size_t)r;
  }
  return (long long)done;
}
long long read_all(int fd, char* p, size_t n, off_t off) {
  size_t done = 0;
  while (done < n) {
    ssize_t r = pread(fd, p + done, n - done, off + (off_t)done);
    if (r < 0) { if (errno == EINTR) continue; return -(long long)errno; }
    if (r == 0) break;  // end of file
    done += (size_t)r;
  }
  return (long long)done;
}

}  // namespace

// Both return the number of bytes transferred (reads: < nbytes at end of file), -errno for I/O errors, -(100000 + cudaError_t) for CUDA errors.
// `ptr` is a device pointer when is_device != 0 (the copies are ordered on `st`, and the call returns after the data is on disk /
// the last host-to-device copy has completed), else a host pointer.
AB_API long long ab_file_write(const char* path, const void* ptr, long long nbytes, long long file_offset, int is_device, int truncate,
                               cudaStream_t st) {
  if (nbytes < 0 || file_offset < 0) return -EINVAL;
  int fd = open(path, O_WRONLY | O_CREAT | (truncate ? O_TRUNC : 0), 0644);
  if (fd < 0) return -(long long)errno;
  long long rc = 0;
  const char* src = static_cast<const char*>(ptr);
  if (!is_device) {
    rc = write_all(fd, src, (size_t)nbytes, (off_t)file_offset);
  } else {
    std::lock_guard<std::mutex> lock(g_stage.mu);
    int e = g_stage.ensure();
    if (e) { close(fd); return -(100000LL + e); }
    const long long n_chunks = (nbytes + (long long)kChunk - 1) / (long long)kChunk;
    auto issue = [&](long long i) -> cudaError_t {
      const size_t len = (size_t)((i + 1 == n_chunks) ? nbytes - i * (long long)kChunk : (long long)kChunk);
      cudaError_t ce = cudaMemcpyAsync(g_stage.buf[i & 1], src + i * (long long)kChunk, len, cudaMemcpyDeviceToHost, st);
      if (ce != cudaSuccess) return ce;
      return cudaEventRecord(g_stage.ev[i & 1], st);
    };
    cudaError_t ce = n_chunks > 0 ? issue(0) : cudaSuccess;
    for (long long i = 0; i < n_chunks && ce == cudaSuccess && rc >= 0; i++) {
      if (i + 1 < n_chunks) ce = issue(i + 1);  // its staging buffer was written out one iteration ago
      if (ce != cudaSuccess) break;
      ce = cudaEventSynchronize(g_stage.ev[i & 1]);
      if (ce != cudaSuccess) break;
      const size_t len = (size_t)((i + 1 == n_chunks) ? nbytes - i * (long long)kChunk : (long long)kChunk);
      long long w = write_all(fd, static_cast<const char*>(g_stage.buf[i & 1]), len, (off_t)(file_offset + i * (long long)kChunk));
      if (w < 0) rc = w; else rc += w;
    }
    if (ce != cudaSuccess || rc < 0) cudaStreamSynchronize(st);  // nothing may still be landing in the staging buffers when the lock drops
    if (ce != cudaSuccess) rc = -(100000LL + (int)ce);
  }
  if (close(fd) != 0 && rc >= 0) rc = -(long long)errno;
  return rc;
}

AB_API long long ab_file_read(const char* path, void* ptr, long long nbytes, long long file_offset, int is_device, cudaStream_t st) {
  if (nbytes < 0 || file_offset < 0) return -EINVAL;
  int fd = open(path, O_RDONLY);
  if (fd < 0) return -(long long)errno;
#ifdef POSIX_FADV_SEQUENTIAL
  posix_fadvise(fd, (off_t)file_offset, (off_t)nbytes, POSIX_FADV_SEQUENTIAL);
#endif
  long long rc = 0;
  char* dst = static_cast<char*>(ptr);
  if (!is_device) {
    rc = read_all(fd, dst, (size_t)nbytes, (off_t)file_offset);
  } else {
    std::lock_guard<std::mutex> lock(g_stage.mu);
    int e = g_stage.ensure();
    if (e) { close(fd); return -(100000LL + e); }
    const long long n_chunks = (nbytes + (long long)kChunk - 1) / (long long)kChunk;
    cudaError_t ce = cudaSuccess;
    bool used[2] = {false, false};
    for (long long i = 0; i < n_chunks; i++) {
      const int b = (int)(i & 1);
      if (used[b]) { ce = cudaEventSynchronize(g_stage.ev[b]); if (ce != cudaSuccess) break; }  // its previous H2D copy has drained
      const size_t len = (size_t)((i + 1 == n_chunks) ? nbytes - i * (long long)kChunk : (long long)kChunk);
      long long r = read_all(fd, static_cast<char*>(g_stage.buf[b]), len, (off_t)(file_offset + i * (long long)kChunk));
      if (r < 0) { rc = r; break; }
      if (r > 0) {
        ce = cudaMemcpyAsync(dst + i * (long long)kChunk, g_stage.buf[b], (size_t)r, cudaMemcpyHostToDevice, st);
        if (ce == cudaSuccess) ce = cudaEventRecord(g_stage.ev[b], st);
        if (ce != cudaSuccess) break;
        used[b] = true;
        rc += r;
      }
      if ((size_t)r < len) break;  // end of file
    }
    for (int b = 0; b < 2; b++)
      if (used[b]) { cudaError_t s = cudaEventSynchronize(g_stage.ev[b]); if (ce == cudaSuccess) ce = s; }
    if (ce != cudaSuccess) rc = -(100000LL + (int)ce);
  }
  close(fd);
  return rc;
}
