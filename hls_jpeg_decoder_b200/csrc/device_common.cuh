// csrc/device_common.cuh -- small device helpers shared by kernels.cu and selfsync.cu.
#pragma once
#include <stdint.h>
#include "hjd_types.h"

// Byte offsets inside HjdHuffTable (lut, limit, delta, vals) for 32-bit shared-window addressing.
#define HJD_TAB_LIMIT_OFF (HJD_LUT_SIZE * 2)
#define HJD_TAB_DELTA_OFF (HJD_LUT_SIZE * 2 + 68)
#define HJD_TAB_VALS_OFF  (HJD_LUT_SIZE * 2 + 136)

__device__ __forceinline__ uint32_t hjd_lds_u16(uint32_t a) { uint16_t v; asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t hjd_lds_u8(uint32_t a) { uint32_t v; asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t hjd_lds_u32(uint32_t a) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void hjd_sts_u16_sync(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "h"((uint16_t)v) : "memory"); }
__device__ __forceinline__ void hjd_sts_v2_sync(uint32_t a, uint32_t x, uint32_t y) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(a), "r"(x), "r"(y) : "memory"); }
__device__ __forceinline__ uint2 hjd_lds_v2_sync(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint4 hjd_lds_v4_sync(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void hjd_sts_zero16_sync(uint32_t a) { asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" :: "r"(a), "r"(0) : "memory"); }
// PTX shifts clamp the amount to 32 (result 0), unlike C++ where a shift by 32 is undefined.
__device__ __forceinline__ uint32_t hjd_shr(uint32_t v, uint32_t n) { uint32_t r; asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n)); return r; }
__device__ __forceinline__ uint32_t hjd_shl(uint32_t v, uint32_t n) { uint32_t r; asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n)); return r; }

