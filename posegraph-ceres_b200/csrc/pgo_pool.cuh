// pgo_pool.cuh -- per-device caches of the host runtime: device memory blocks, streams, events and
// small pinned buffers survive pgo_graph_destroy and are handed to the next graph, so that the
// one-shot entry point (pgo_solve_pose_graph: upload + solve + download, what ceres::Solve is for the
// reference) does not pay cudaMalloc / cudaFree / cudaMallocHost / cudaStreamCreate on every call.
// pgo_release_cached_memory() returns everything to the driver.
#pragma once

#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <vector>

namespace pgo {

struct DevicePool {
  std::mutex mu;
  std::multimap<size_t, void*> free_blocks;     // size -> device pointer
  std::map<void*, size_t> alloc_size;           // true (rounded) size of every block this pool ever handed out
  size_t cached_bytes = 0;
  std::vector<cudaStream_t> streams;
  std::vector<cudaEvent_t> events;
  std::vector<void*> pinned;                    // kPinnedBytes each
  int num_sms = 0;
  // per-device facts about kernels: occupancy results and "function attribute already set on THIS device" flags
  // (cudaFuncSetAttribute applies to the current device only, so a process-wide static would be wrong)
  int cache_val[64] = {};
  bool cache_has[64] = {};
};

enum PoolCacheKey {
  kCachePcgPerSm = 0, kCachePcgCluster = 1, kCacheCholPerSm = 2, kCacheCholCluster = 3, kCacheAmgDenseAttr = 4,
  kCacheLinAttrBase = 8   // + variant index (< 16)
};

constexpr size_t kPinnedBytes = 4096;
constexpr int kMaxDevices = 64;

inline DevicePool& device_pool(int device) {
  static DevicePool pools[kMaxDevices];
  return pools[device < 0 || device >= kMaxDevices ? 0 : device];
}

inline bool pool_cache_get(int device, int key, int* v) {
  DevicePool& P = device_pool(device);
  std::lock_guard<std::mutex> lk(P.mu);
  if (!P.cache_has[key]) return false;
  *v = P.cache_val[key];
  return true;
}
inline void pool_cache_set(int device, int key, int v) {
  DevicePool& P = device_pool(device);
  std::lock_guard<std::mutex> lk(P.mu);
  P.cache_val[key] = v; P.cache_has[key] = true;
}

inline size_t pool_round(size_t bytes) {
  if (bytes < 512) return 512;
  return (bytes + 511) & ~(size_t)511;
}

// Device memory: exact-ish fit from the cache (<= 25% slack), else cudaMalloc (flushing the cache on failure).
inline cudaError_t pool_alloc(int device, void** out, size_t bytes) {
  DevicePool& P = device_pool(device);
  const size_t want = pool_round(bytes);
  {
    std::lock_guard<std::mutex> lk(P.mu);
    auto it = P.free_blocks.lower_bound(want);
    if (it != P.free_blocks.end() && it->first <= want + want / 4) {
      *out = it->second;   // (its true size stays registered in alloc_size)
      P.cached_bytes -= it->first;
      P.free_blocks.erase(it);
      return cudaSuccess;
    }
  }
  cudaError_t e = cudaMalloc(out, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    std::vector<void*> drop;
    {
      std::lock_guard<std::mutex> lk(P.mu);
      for (auto& kv : P.free_blocks) { drop.push_back(kv.second); P.alloc_size.erase(kv.second); }
      P.free_blocks.clear();
      P.cached_bytes = 0;
    }
    for (void* p : drop) cudaFree(p);
    e = cudaMalloc(out, want);
  }
  if (e == cudaSuccess) {
    std::lock_guard<std::mutex> lk(P.mu);
    P.alloc_size[*out] = want;
  }
  return e;
}

// The caller guarantees no work that touches the block is still in flight.
inline void pool_free(int device, void* p, size_t bytes) {
  if (!p) return;
  DevicePool& P = device_pool(device);
  size_t sz = pool_round(bytes);
  bool keep = true;
  {
    std::lock_guard<std::mutex> lk(P.mu);
    // a reused block may be larger than what its last user asked for: file it under its TRUE size
    auto it = P.alloc_size.find(p);
    if (it != P.alloc_size.end()) sz = it->second;
    if (P.cached_bytes + sz > ((size_t)8 << 30)) keep = false;   // never sit on more than 8 GiB
    if (keep) { P.free_blocks.emplace(sz, p); P.cached_bytes += sz; }
    else P.alloc_size.erase(p);
  }
  if (!keep) cudaFree(p);
}

inline cudaError_t pool_stream(int device, cudaStream_t* s) {
  DevicePool& P = device_pool(device);
  {
    std::lock_guard<std::mutex> lk(P.mu);
    if (!P.streams.empty()) { *s = P.streams.back(); P.streams.pop_back(); return cudaSuccess; }
  }
  return cudaStreamCreateWithFlags(s, cudaStreamNonBlocking);
}
inline void pool_stream_release(int device, cudaStream_t s) {
  if (!s) return;
  DevicePool& P = device_pool(device);
  std::lock_guard<std::mutex> lk(P.mu);
  P.streams.push_back(s);
}
inline cudaError_t pool_event(int device, cudaEvent_t* ev) {
  DevicePool& P = device_pool(device);
  {
    std::lock_guard<std::mutex> lk(P.mu);
    if (!P.events.empty()) { *ev = P.events.back(); P.events.pop_back(); return cudaSuccess; }
  }
  return cudaEventCreate(ev);
}
inline void pool_event_release(int device, cudaEvent_t ev) {
  if (!ev) return;
  DevicePool& P = device_pool(device);
  std::lock_guard<std::mutex> lk(P.mu);
  P.events.push_back(ev);
}
inline cudaError_t pool_pinned(int device, void** p) {
  DevicePool& P = device_pool(device);
  {
    std::lock_guard<std::mutex> lk(P.mu);
    if (!P.pinned.empty()) { *p = P.pinned.back(); P.pinned.pop_back(); return cudaSuccess; }
  }
  return cudaMallocHost(p, kPinnedBytes);
}
inline void pool_pinned_release(int device, void* p) {
  if (!p) return;
  DevicePool& P = device_pool(device);
  std::lock_guard<std::mutex> lk(P.mu);
  P.pinned.push_back(p);
}
inline int pool_num_sms(int device) {
  DevicePool& P = device_pool(device);
  if (P.num_sms == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) { cudaGetLastError(); n = 1; }
    P.num_sms = n;
  }
  return P.num_sms;
}

// Return every cached resource of `device` (or of all devices when device < 0) to the driver.
inline void pool_release(int device) {
  for (int d = 0; d < kMaxDevices; ++d) {
    if (device >= 0 && d != device) continue;
    DevicePool& P = device_pool(d);
    std::vector<void*> blocks, pins;
    std::vector<cudaStream_t> st;
    std::vector<cudaEvent_t> ev;
    {
      std::lock_guard<std::mutex> lk(P.mu);
      for (auto& kv : P.free_blocks) { blocks.push_back(kv.second); P.alloc_size.erase(kv.second); }
      P.free_blocks.clear(); P.cached_bytes = 0;
      st.swap(P.streams); ev.swap(P.events); pins.swap(P.pinned);
    }
    if (blocks.empty() && st.empty() && ev.empty() && pins.empty()) continue;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(d);
    for (void* p : blocks) cudaFree(p);
    for (void* p : pins) cudaFreeHost(p);
    for (cudaStream_t s : st) cudaStreamDestroy(s);
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
    cudaSetDevice(prev);
  }
}

}  // namespace pgo
