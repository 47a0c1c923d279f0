// Shared helpers for the sm_100a kernels of the FRCNN extraction path.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace vltk {

typedef __nv_bfloat16 bf16;

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

// ---- error plumbing: nothing throws across the C ABI -------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define VLTK_CUDA(call)                                                                  \
  do {                                                                                   \
    cudaError_t _e = (call);                                                             \
    if (_e != cudaSuccess) {                                                             \
      vltk::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      return -1;                                                                         \
    }                                                                                    \
  } while (0)

#define VLTK_CHECK(cond, ...)            \
  do {                                   \
    if (!(cond)) {                       \
      vltk::set_error(__VA_ARGS__);      \
      return -2;                         \
    }                                    \
  } while (0)

#define VLTK_LAUNCH_CHECK() VLTK_CUDA(cudaGetLastError())

// cudaFuncSetAttribute (the >48 KB dynamic shared memory opt-in) is PER DEVICE: a process that drives several
// GPUs must repeat it on each one.  Returns true the first time it is called for the current device with a
// given per-kernel flag array (one `static DeviceOnce` per kernel / template instantiation).
struct DeviceOnce {
  bool done[64] = {};
  bool first() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return true;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};

// ---- element access helpers (fp32 or bf16 activations) -----------------------------------
__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const bf16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(bf16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// order-preserving float <-> uint key (ascending)
__host__ __device__ __forceinline__ uint32_t float_to_key(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

}  // namespace vltk
