// device_util.cuh -- CUDA error checking and RAII device buffers for the context.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

#include "../error.hpp"

namespace pecs {

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
  if (e != cudaSuccess)
    throw StatusError(PECS_ERR_CUDA, std::string(what) + " failed: " + cudaGetErrorString(e) + " (" + file + ":" +
                                         std::to_string(line) + ")");
}
#define PECS_CUDA(call) ::pecs::cuda_check((call), #call, __FILE__, __LINE__)

template <class T>
class DeviceBuffer {
public:
  DeviceBuffer() = default;
  explicit DeviceBuffer(size_t n) { resize(n); }
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
  DeviceBuffer(DeviceBuffer&& o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr; o.n_ = 0; }
  DeviceBuffer& operator=(DeviceBuffer&& o) noexcept {
    if (this != &o) {
      release();
      p_ = o.p_;
      n_ = o.n_;
      o.p_ = nullptr;
      o.n_ = 0;
    }
    return *this;
  }
  ~DeviceBuffer() { release(); }
  void resize(size_t n) {
    release();
    n_ = n;
    if (n) PECS_CUDA(cudaMalloc(&p_, n * sizeof(T)));
  }
  void upload(const T* h, size_t n) {
    if (n != n_) resize(n);
    if (n) PECS_CUDA(cudaMemcpy(p_, h, n * sizeof(T), cudaMemcpyHostToDevice));
  }
  void upload(const std::vector<T>& h) { upload(h.data(), h.size()); }
  void zero() {
    if (n_) PECS_CUDA(cudaMemset(p_, 0, n_ * sizeof(T)));
  }
  T* get() const { return p_; }
  size_t size() const { return n_; }
  size_t bytes() const { return n_ * sizeof(T); }

private:
  void release() {
    if (p_) cudaFree(p_);
    p_ = nullptr;
    n_ = 0;
  }
  T* p_ = nullptr;
  size_t n_ = 0;
};

#ifdef __CUDACC__
// Loads of STEP-VARYING vectors (states, right-hand sides, work vectors of the solves, child-update buffers).
// Most kernels of a step are launched programmatically (launch_pdl): a consumer grid is resident -- prologue running,
// L1 of its SMs alive -- while its producer still writes these vectors.  PTX allows ld.global.nc (__ldg, or what nvcc
// emits for `const T* __restrict__`) only for data that is read-only for the WHOLE lifetime of the grid; round 1 read
// these vectors that way, outside the contract.  Two coherent forms replace it:
//   ld_vec  : plain ld.global (L1-cached, coherent at grid-dependency boundaries: griddepcontrol.wait / kernel start
//             make the producer's stores visible to it).  For kernels whose producers are whole GRIDS: the ELL
//             mat-vecs (their gathers of x re-use lines across a warp: L1 matters, 53 -> 63 us per kernel without it)
//             and the assembly kernels.
//   ld_step : ld.global.cg (L2, the point of coherence).  For the level kernels of the solves, whose producers are
//             other thread blocks of the SAME or of a concurrently running grid, ordered by per-front release/acquire
//             counters (solve_kernels.cu): a line another block of this SM pulled into L1 earlier must never serve
//             them.  Nothing they read this way is re-used from L1 anyway.
// .nc stays for the static tables only.  -DPECS_B200_NC_STEP_VECTORS=1 restores the round-1 loads (the A side of the
// A/B experiment in scripts/race_repro.py; never shipped).
#ifndef PECS_B200_NC_STEP_VECTORS
#define PECS_B200_NC_STEP_VECTORS 0
#endif
template <class T>
__device__ __forceinline__ T ld_step(const T* p) {
#if PECS_B200_NC_STEP_VECTORS
  return __ldg(p);
#else
  return __ldcg(p);
#endif
}
template <class T>
__device__ __forceinline__ T ld_vec(const T* p) {
#if PECS_B200_NC_STEP_VECTORS
  return __ldg(p);
#else
  return *p; // the callers hold these pointers without __restrict__, so nvcc emits a plain ld.global
#endif
}
// 256-bit form (sm_100a): one instruction per 4-vector of nodal values
__device__ __forceinline__ void ld_vec4(const double* p, double v[4]) {
#if PECS_B200_NC_STEP_VECTORS
  asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
#else
  asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p) : "memory");
#endif
}

// Launch with programmatic stream serialization: the kernel may start while its predecessor in the stream is still
// running and synchronises itself with griddepcontrol.wait (wait_for_predecessor).  PECS_B200_PDL=0 launches plainly.
template <class... KArgs, class... Args>
inline void launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t s, Args... args) {
  static const bool pdl = [] {
    const char* e = std::getenv("PECS_B200_PDL");
    return !(e && e[0] == '0');
  }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}


// Device side of a programmatic dependent launch: everything before this call overlaps the tail of the previous
// kernel in the stream (launch latency, descriptor loads, prefetches of static tables); after it the predecessor's
// results are visible.  The successor is released right away: it may start ITS prologue while this kernel computes.
__device__ __forceinline__ void wait_for_predecessor() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#endif

} // namespace pecs
