// csrc/device_common.cuh -- small device helpers shared by kernels.cu and selfsync.cu.
#pragma once
#include <stdint.h>
#include "hjd_types.h"

// Byte offset of the second-level table inside HjdHuffTable, for 32-bit shared-window addressing.
#define HJD_TAB_LUT2_OFF (HJD_LUT_SIZE * 2)

// Meaning of a decoded symbol (see HjdHuffTable); also used by the host when it fills the LUT.
__host__ __device__ __forceinline__ uint32_t hjd_sym_fields(uint32_t len, uint32_t sym, bool is_ac)
{
    const uint32_t size = sym & 15u, run = sym >> 4;
    if (!is_ac) return HJD_SYM_FIELDS(len, size, 1);                          // DC: loadjpg.cpp:616-667
    if (size) return HJD_SYM_FIELDS(len, size, run + 1);                      // loadjpg.cpp:778-806
    return HJD_SYM_FIELDS(len, 0, run == 0 ? 63 : (run == 15 ? 16 : 0));      // EOB / ZRL / ignored, 771-775
}

__device__ __forceinline__ uint32_t hjd_lds_u16(uint32_t a) { uint16_t v; asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t hjd_lds_u8(uint32_t a) { uint32_t v; asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t hjd_lds_u32(uint32_t a) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t hjd_lds_u16_sync(uint32_t a) { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void hjd_sts_u16_sync(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "h"((uint16_t)v) : "memory"); }
__device__ __forceinline__ void hjd_sts_v2_sync(uint32_t a, uint32_t x, uint32_t y) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(a), "r"(x), "r"(y) : "memory"); }
__device__ __forceinline__ uint2 hjd_lds_v2_sync(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint4 hjd_lds_v4_sync(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void hjd_sts_zero16_sync(uint32_t a) { asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" :: "r"(a), "r"(0) : "memory"); }
// PTX shifts clamp the amount to 32 (result 0), unlike C++ where a shift by 32 is undefined.
__device__ __forceinline__ uint32_t hjd_shr(uint32_t v, uint32_t n) { uint32_t r; asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n)); return r; }
__device__ __forceinline__ uint32_t hjd_shl(uint32_t v, uint32_t n) { uint32_t r; asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n)); return r; }


// The symbol whose code starts the 32-bit window `hi` (MSB first): len | size << 5 | advance << 9 (see
// HjdHuffTable).  t = shared-window address of the table.  Codes longer than the first-level table take
// one more load; an unassigned code yields HJD_BAD_ENTRY (advance 127).
__device__ __forceinline__ uint32_t hjd_lookup(uint32_t t, uint32_t hi)
{
    uint32_t e = hjd_lds_u16(t + ((hi >> (32 - HJD_LUT_BITS)) << 1));
    if ((e & 31u) == 0) {
        const uint32_t idx = hjd_shr((hi >> 16) & ((1u << (16 - HJD_LUT_BITS)) - 1u), (e >> 5) & 7u);
        e = hjd_lds_u16(t + HJD_TAB_LUT2_OFF + ((e >> 8) << 2) + (idx << 1));
    }
    return e;
}
