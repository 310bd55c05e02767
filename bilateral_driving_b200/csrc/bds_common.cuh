// Shared helpers for the sm_100a kernels of libbds_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/bds.h"

#define BDS_HD __host__ __device__ __forceinline__
#define BDS_D __device__ __forceinline__

namespace bds {

void set_error(const char* fmt, ...);

#define BDS_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t err__ = (expr);                                                           \
    if (err__ != cudaSuccess) {                                                           \
      bds::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(err__)); \
      return -2;                                                                          \
    }                                                                                     \
  } while (0)

#define BDS_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      bds::set_error(__VA_ARGS__);    \
      return -1;                      \
    }                                 \
  } while (0)

extern unsigned long long g_launches;  // kernels launched by this library (statistics only; relaxed atomic:
                                       // the entry points are re-entrant across host threads / streams)
}  // namespace bds
#define BDS_CHECK_LAUNCH()      \
  do {                          \
    __atomic_fetch_add(&bds::g_launches, 1ull, __ATOMIC_RELAXED); \
    BDS_CHECK_CUDA(cudaGetLastError()); \
  } while (0)
namespace bds {

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

constexpr int kTile = BDS_TILE;
constexpr float kAlphaMin = 1.0f / 255.0f;
constexpr float kAlphaMax = 0.999f;
constexpr float kTStop = 1e-4f;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kLog2_255 = 7.994353436858858f;   // alpha >= 1/255  <=>  log2(alpha) >= -kLog2_255

#ifdef __CUDACC__
// warp helpers ---------------------------------------------------------------------------------
BDS_D float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

BDS_D unsigned lane_id() {
  unsigned l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}

// streaming (read-once) 128-bit load / store that do not allocate in L1
BDS_D float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
BDS_D void st_stream_f4(float4* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w));
}

// Packed fp32x2 arithmetic (sm_100a FFMA2 / FADD2 / FMUL2): two IEEE fp32 operations per issue slot; a scalar
// operand is broadcast by the instruction itself.  Same rounding as the scalar forms.
typedef unsigned long long f32x2;
BDS_D f32x2 pk2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
BDS_D void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
BDS_D f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
BDS_D void fma2_acc(f32x2& c, f32x2 a, f32x2 b) {   // c += a * b, accumulator updated in place
  asm("fma.rn.ftz.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
}
BDS_D f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
BDS_D f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
BDS_D f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("sub.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// index of the highest set bit (FLO.U32); x != 0
BDS_D int bfind_u32(unsigned x) {
  int r;
  asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
  return r;
}

// 1u << pos (BMSK: no constant register needed); pos in [0, 31]
BDS_D unsigned bit_mask(int pos) {
  unsigned r;
  asm("bmsk.clamp.b32 %0, %1, 1;" : "=r"(r) : "r"(pos));
  return r;
}
// 32-bit shared-window address of a shared-memory pointer, and back
BDS_D uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <typename T>
BDS_D const T* smem_ptr(uint32_t a) { return reinterpret_cast<const T*>(__cvta_shared_to_generic((size_t)a)); }

// fire-and-forget fp32 reduction into global memory (RED.ADD.F32)
BDS_D void red_add(float* p, float v) { atomicAdd(p, v); }
// four consecutive floats, 16-byte aligned, in one reduction (RED.E.ADD.F32x4)
BDS_D void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
#endif  // __CUDACC__

}  // namespace bds
