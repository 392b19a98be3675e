// TEST INFRASTRUCTURE - NOT PRODUCT CODE, NOT A CPU FALLBACK.
// A minimal stand-in for <cuda_runtime.h> that lets the device sources of openifem_b200/csrc be compiled with g++ and their
// kernels be EXECUTED on the CPU by a small SIMT emulator (cpu_emul_engine.cpp): every CUDA thread of a block is a fiber,
// __syncthreads / __syncwarp / __shfl_xor_sync are real barriers between fibers, blocks run one after the other. It exists
// so that the `-m gpu` parity tests can be replayed on a machine without a GPU (tests/test_emulated_device_cpu.py) - the
// launch configurations, the indexing, the shared-memory staging and the host orchestration of the product sources run
// unchanged. The product library (openifem_b200/lib/libopenifem_b200.so, built by nvcc for sm_100a) never sees this header;
// the emulated build is a different file (tests/cpu_emul/_build/libopenifem_b200_cpuemul.so) that only the test harness loads.
#pragma once
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define IFEM_CPU_EMULATION 1

struct dim3
{
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
inline int2 make_int2(int x, int y) { return {x, y}; }
inline float2 make_float2(float x, float y) { return {x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }

namespace cpu_emul
{
  struct Idx { unsigned x, y, z; };
  extern Idx g_tid, g_bid;
  extern dim3 g_block, g_grid;
  void launch(dim3 grid, dim3 block, size_t dyn_smem_bytes, bool uses_barriers, const char *what, const std::function<void()> &body);
  void *dyn_smem();
  void sync_threads();
  void sync_warp();
  uint64_t shfl_xor_bits(uint64_t bits, int lane_mask);
  extern long long g_launches, g_fiber_launches;
} // namespace cpu_emul

#define threadIdx (cpu_emul::g_tid)
#define blockIdx (cpu_emul::g_bid)
#define blockDim (cpu_emul::g_block)
#define gridDim (cpu_emul::g_grid)

inline void __syncthreads() { cpu_emul::sync_threads(); }
inline void __syncwarp(unsigned = 0xffffffffu) { cpu_emul::sync_warp(); }
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask, int = 32)
{
  static_assert(sizeof(T) <= 8, "shuffle of up to 8 bytes");
  uint64_t bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  bits = cpu_emul::shfl_xor_bits(bits, lane_mask);
  T r;
  std::memcpy(&r, &bits, sizeof(T));
  return r;
}
template <typename T> inline T __ldcs(const T *p) { return *p; }
template <typename T> inline T __ldg(const T *p) { return *p; }
template <typename T> inline T __ldcg(const T *p) { return *p; }
template <typename T> inline void __stcg(T *p, T v) { *p = v; }
inline void __threadfence() {}
inline void __threadfence_system() {}
inline long long __double_as_longlong(double v) { long long r; std::memcpy(&r, &v, 8); return r; }
inline double __longlong_as_double(long long v) { double r; std::memcpy(&r, &v, 8); return r; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
// one host thread runs all fibers: plain read-modify-write is atomic here
template <typename T, typename U> inline T atomicAdd(T *p, U v) { const T old = *p; *p = old + (T)v; return old; }
template <typename T, typename U> inline T atomicExch(T *p, U v) { const T old = *p; *p = (T)v; return old; }

// ---- runtime API: host memory stands in for device memory, everything is synchronous ----
typedef int cudaError_t;
enum { cudaSuccess = 0 };
typedef struct cpu_emul_stream *cudaStream_t;
typedef struct cpu_emul_event { std::chrono::steady_clock::time_point t; } *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyHostToHost };
enum { cudaStreamNonBlocking = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { int multiProcessorCount; char name[64]; size_t totalGlobalMem; };
inline const char *cudaGetErrorName(cudaError_t) { return "cpu_emul"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { p->multiProcessorCount = 4; std::strcpy(p->name, "cpu_emul"); p->totalGlobalMem = 1ull << 34; return cudaSuccess; }
// "device" allocations carry 256-byte red zones on both sides, filled with a pattern that cudaFree and a process-exit hook
// verify: a kernel that writes just outside a buffer is reported with the size of the buffer (a poor man's memcheck)
namespace cpu_emul
{
  void *guarded_alloc(size_t n);
  void guarded_free(void *p);
} // namespace cpu_emul
template <typename T> inline cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)cpu_emul::guarded_alloc(n); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFree(void *p) { cpu_emul::guarded_free(p); return cudaSuccess; }
template <typename T> inline cudaError_t cudaMallocHost(T **p, size_t n) { *p = (T *)std::malloc(n ? n : 1); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void *p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new cpu_emul_event; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess; }
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
// no peer memory between emulated ranks: the product falls back to its communicator path (peer.h)
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *, void *) { return 1; }
inline cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned) { return 1; }
inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }

// CUDA puts these in the global namespace
template <typename T> inline T min(T a, T b) { return b < a ? b : a; }
template <typename T> inline T max(T a, T b) { return a < b ? b : a; }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
using std::isfinite;
using std::isnan;
using std::isinf;
