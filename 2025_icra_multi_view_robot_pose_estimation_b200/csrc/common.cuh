// common.cuh — device helpers shared by the mvgeo kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "mvgeo.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "mvgeo kernels are written for sm_100a (B200); build with -gencode arch=compute_100a,code=sm_100a"
#endif

#define MVGEO_CHECK_LAUNCH()                          \
  do {                                                \
    cudaError_t e__ = cudaGetLastError();             \
    if (e__ != cudaSuccess) return (int)e__;          \
  } while (0)

#define MVGEO_CUDA(call)                              \
  do {                                                \
    cudaError_t e__ = (call);                         \
    if (e__ != cudaSuccess) return (int)e__;          \
  } while (0)

namespace mvgeo {

constexpr float kLog2e = 1.4426950408889634f;
// Soft-arg-max skip threshold (natural-log units): elements whose weight relative to the
// peak is below exp(-kSoftSkip) = 1.3e-14 are not accumulated (worst-case centroid error
// n_pixels * W * 1.3e-14 < 3e-6 px at 480x640).
constexpr float kSoftSkip = 32.0f;

__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float max_nan_f32(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ uint32_t max_nan_bf16x2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("max.NaN.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t max_nan_f16x2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("max.NaN.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// ---- packed f32x2 arithmetic (sm_100: one FFMA2 / FADD2 / FMUL2 issue slot for two lanes) ----
typedef uint64_t f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float sum2(f32x2 v) {
  float lo, hi;
  unpack2(v, lo, hi);
  return lo + hi;
}

// ---- mbarrier + TMA bulk-copy primitives (async proxy; SASS: SYNCS.*, UBLKCP) -------------
// Shared-memory operands are 32-bit shared-window addresses computed once (smem_u32) and then
// advanced with integer arithmetic, so no generic->shared conversion sits in the streaming loop.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float lds32f(uint32_t addr) {
  float r;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(addr));
  return r;
}
__device__ __forceinline__ uint32_t lds32u(uint32_t addr) {
  uint32_t r;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr));
  return r;
}
__device__ __forceinline__ void sts32f(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MVGEO_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra MVGEO_DONE_%=;\n"
      "bra MVGEO_WAIT_%=;\n"
      "MVGEO_DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity), "r"(0x989680u)  // suspend-time hint: sleep in hardware instead of re-polling
      : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on `bar` (16-byte aligned, size % 16 == 0).
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst_smem, const void* src_gmem, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src_gmem), "r"(bytes), "r"(bar)
               : "memory");
}

// (value, index) ordering of torch.argmax: larger value wins, NaN is maximal, ties (equal
// values, -0 == +0, or two NaNs) go to the lower index. True when (v2,i2) beats (v1,i1).
__device__ __forceinline__ bool argmax_better(float v1, int i1, float v2, int i2) {
  const bool n1 = (v1 != v1), n2 = (v2 != v2);
  if (n1 != n2) return n2;
  if (!n1) {
    if (v2 > v1) return true;
    if (v2 < v1) return false;
  }
  return i2 < i1;
}

__device__ __forceinline__ void warp_argmax(float& v, int& i) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, off);
    const int oi = __shfl_xor_sync(0xffffffffu, i, off);
    if (argmax_better(v, i, ov, oi)) {
      v = ov;
      i = oi;
    }
  }
}

// Fixed-order (deterministic) warp sum.
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// 2^t for two elements on the FMA / ALU pipes (no MUFU): t is clamped to >= -126 (the result is then 2^-126, not 0),
// split by the magic-number add into n = round(t) and f = t - n in [-0.5, 0.5], 2^f by a degree-5 minimax polynomial
// (relative error 2.2e-7 evaluated in f32, the same as ex2.approx), and n added to the exponent field. t <= 127.
__device__ __forceinline__ f32x2 exp2_poly2(f32x2 t) {
  float lo, hi;
  unpack2(t, lo, hi);
  t = pack2(fmaxf(lo, -126.0f), fmaxf(hi, -126.0f));
  const f32x2 magic = pack2(12582912.0f, 12582912.0f);  // 1.5 * 2^23
  const f32x2 r = add2(t, magic);
  const f32x2 f = sub2(t, sub2(r, magic));
  f32x2 q = pack2(1.327637467e-03f, 1.327637467e-03f);
  q = fma2(q, f, pack2(9.675514884e-03f, 9.675514884e-03f));
  q = fma2(q, f, pack2(5.550713465e-02f, 5.550713465e-02f));
  q = fma2(q, f, pack2(2.402212024e-01f, 2.402212024e-01f));
  q = fma2(q, f, pack2(6.931469440e-01f, 6.931469440e-01f));
  q = fma2(q, f, pack2(1.000000119e+00f, 1.000000119e+00f));
  float rl, rh;
  unpack2(r, rl, rh);
  unpack2(q, lo, hi);
  // the low mantissa bits of r hold n (two's complement); << 23 moves them onto the exponent field
  lo = __uint_as_float(__float_as_uint(lo) + (__float_as_uint(rl) << 23));
  hi = __uint_as_float(__float_as_uint(hi) + (__float_as_uint(rh) << 23));
  return pack2(lo, hi);
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// Element access for the three belief-map dtypes. A "chunk" is one 16-byte vector.
template <int DT> struct Elem;

template <> struct Elem<MVGEO_F32> {
  static constexpr int kBytes = 4;
  static constexpr int kPerChunk = 4;
  using carrier = float;  // how a chunk maximum is kept in shared memory
  // elements (2i, 2i+1) of a chunk as two floats (already a register pair for f32 maps)
  __device__ static __forceinline__ void pair(const uint4& c, int i, float& lo, float& hi) {
    lo = __uint_as_float(i == 0 ? c.x : c.z);
    hi = __uint_as_float(i == 0 ? c.y : c.w);
  }
  __device__ static __forceinline__ float chunk_max(const uint4& c) {
    return max_nan_f32(max_nan_f32(__uint_as_float(c.x), __uint_as_float(c.y)),
                       max_nan_f32(__uint_as_float(c.z), __uint_as_float(c.w)));
  }
  __device__ static __forceinline__ float get(const uint4& c, int j) {
    const uint32_t w = j == 0 ? c.x : j == 1 ? c.y : j == 2 ? c.z : c.w;
    return __uint_as_float(w);
  }
  // "vertical" maximum of a chunk kept in a 32-bit register (for f32 simply the chunk maximum),
  // merge of two such values, and the final horizontal step giving the float maximum
  __device__ static __forceinline__ uint32_t vmax(const uint4& c) { return __float_as_uint(chunk_max(c)); }
  __device__ static __forceinline__ uint32_t vmerge(uint32_t a, uint32_t b) {
    return __float_as_uint(max_nan_f32(__uint_as_float(a), __uint_as_float(b)));
  }
  __device__ static __forceinline__ float vfinish(uint32_t v) { return __uint_as_float(v); }
  __device__ static __forceinline__ carrier pack(float m) { return m; }
  __device__ static __forceinline__ float unpack(carrier m) { return m; }
  __device__ static __forceinline__ float load(const void* base, int64_t i) {
    return __ldg(reinterpret_cast<const float*>(base) + i);
  }
  __device__ static __forceinline__ void store(void* base, int64_t i, float v) {
    reinterpret_cast<float*>(base)[i] = v;
  }
  __device__ static __forceinline__ uint4 neg_inf_chunk() {
    return make_uint4(0xff800000u, 0xff800000u, 0xff800000u, 0xff800000u);
  }
};

template <> struct Elem<MVGEO_BF16> {
  static constexpr int kBytes = 2;
  static constexpr int kPerChunk = 8;
  using carrier = uint16_t;
  __device__ static __forceinline__ void pair(const uint4& c, int i, float& lo, float& hi) {
    const uint32_t w = i == 0 ? c.x : i == 1 ? c.y : i == 2 ? c.z : c.w;
    lo = __uint_as_float(w << 16);
    hi = __uint_as_float(w & 0xffff0000u);
  }
  __device__ static __forceinline__ float chunk_max(const uint4& c) {
    const uint32_t m = max_nan_bf16x2(max_nan_bf16x2(c.x, c.y), max_nan_bf16x2(c.z, c.w));
    return max_nan_f32(__uint_as_float(m << 16), __uint_as_float(m & 0xffff0000u));
  }
  __device__ static __forceinline__ float get(const uint4& c, int j) {
    const uint32_t w = (j >> 1) == 0 ? c.x : (j >> 1) == 1 ? c.y : (j >> 1) == 2 ? c.z : c.w;
    return __uint_as_float((j & 1) ? (w & 0xffff0000u) : (w << 16));
  }
  // packed bf16x2 "vertical" maximum (2 SIMD instructions per chunk); the horizontal step is
  // done once per slice: after it both halves hold the maximum and the high half IS the float
  __device__ static __forceinline__ uint32_t vmax(const uint4& c) {
    return max_nan_bf16x2(max_nan_bf16x2(c.x, c.y), max_nan_bf16x2(c.z, c.w));
  }
  __device__ static __forceinline__ uint32_t vmerge(uint32_t a, uint32_t b) { return max_nan_bf16x2(a, b); }
  __device__ static __forceinline__ float vfinish(uint32_t v) {
    const uint32_t m2 = max_nan_bf16x2(v, __byte_perm(v, v, 0x1032));
    return __uint_as_float(m2 & 0xffff0000u);
  }
  __device__ static __forceinline__ carrier pack(float m) { return (uint16_t)(__float_as_uint(m) >> 16); }
  __device__ static __forceinline__ float unpack(carrier m) { return __uint_as_float(((uint32_t)m) << 16); }
  __device__ static __forceinline__ float load(const void* base, int64_t i) {
    return __uint_as_float(((uint32_t)__ldg(reinterpret_cast<const uint16_t*>(base) + i)) << 16);
  }
  __device__ static __forceinline__ void store(void* base, int64_t i, float v) {
    reinterpret_cast<__nv_bfloat16*>(base)[i] = __float2bfloat16_rn(v);
  }
  __device__ static __forceinline__ uint4 neg_inf_chunk() {
    return make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);
  }
};

template <> struct Elem<MVGEO_F16> {
  static constexpr int kBytes = 2;
  static constexpr int kPerChunk = 8;
  using carrier = uint16_t;
  __device__ static __forceinline__ float h2f(uint16_t b) { return __half2float(__ushort_as_half(b)); }
  __device__ static __forceinline__ void pair(const uint4& c, int i, float& lo, float& hi) {
    const uint32_t w = i == 0 ? c.x : i == 1 ? c.y : i == 2 ? c.z : c.w;
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
    lo = f.x;
    hi = f.y;
  }
  __device__ static __forceinline__ float chunk_max(const uint4& c) {
    const uint32_t m = max_nan_f16x2(max_nan_f16x2(c.x, c.y), max_nan_f16x2(c.z, c.w));
    return max_nan_f32(h2f((uint16_t)(m & 0xffffu)), h2f((uint16_t)(m >> 16)));
  }
  __device__ static __forceinline__ float get(const uint4& c, int j) {
    const uint32_t w = (j >> 1) == 0 ? c.x : (j >> 1) == 1 ? c.y : (j >> 1) == 2 ? c.z : c.w;
    return h2f((uint16_t)((j & 1) ? (w >> 16) : (w & 0xffffu)));
  }
  __device__ static __forceinline__ uint32_t vmax(const uint4& c) {
    return max_nan_f16x2(max_nan_f16x2(c.x, c.y), max_nan_f16x2(c.z, c.w));
  }
  __device__ static __forceinline__ uint32_t vmerge(uint32_t a, uint32_t b) { return max_nan_f16x2(a, b); }
  __device__ static __forceinline__ float vfinish(uint32_t v) {
    const uint32_t m2 = max_nan_f16x2(v, __byte_perm(v, v, 0x1032));
    return h2f((uint16_t)m2);
  }
  // every half is exactly representable in float and the maximum IS one of the inputs,
  // so the round trip through __float2half_rn is exact
  __device__ static __forceinline__ carrier pack(float m) { return __half_as_ushort(__float2half_rn(m)); }
  __device__ static __forceinline__ float unpack(carrier m) { return h2f(m); }
  __device__ static __forceinline__ float load(const void* base, int64_t i) {
    return h2f(__ldg(reinterpret_cast<const uint16_t*>(base) + i));
  }
  __device__ static __forceinline__ void store(void* base, int64_t i, float v) {
    reinterpret_cast<__half*>(base)[i] = __float2half_rn(v);
  }
  __device__ static __forceinline__ uint4 neg_inf_chunk() {
    return make_uint4(0xfc00fc00u, 0xfc00fc00u, 0xfc00fc00u, 0xfc00fc00u);
  }
};

}  // namespace mvgeo
