// tools/exp_idct_lanes.cu -- EXPERIMENT, not part of the product.
//
// Question: the shipped IDCT keeps one 8x8 block per thread (64 coefficients in registers, 128 registers,
// 16 warps per SM).  Would a warp-cooperative layout -- 4 lanes per block, two coefficient rows per lane,
// rows / columns exchanged through shared memory, ~1/4 of the registers, twice the resident warps -- be
// faster, as north_star's kernel (2) suggests?
//
// Both kernels do the same arithmetic (de-quantise with C(u)C(v), two passes of FFMA2 chains with the cos
// table in constant memory, truncate, +128, clamp, pack) on the same dense coefficient slab and write 64
// bytes per block.  To isolate the layout, BOTH leave out what the product's kernel also does and what the
// 4-lane layout would have to pay extra for: the zig-zag permutation (free at compile time with one block per
// thread; a 16 x LDS.U16 gather per lane otherwise) and the exact re-evaluation of near-integer samples (all
// 64 terms in one thread; a cross-lane gather otherwise).  So the 4-lane number below is an upper bound of
// what that layout can deliver.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo tools/exp_idct_lanes.cu -o tune/exp_idct_lanes
//   tune/exp_idct_lanes [blocks = 8388608]
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

__constant__ float2 c_cos2[64];     // (cos[x][u], cos[x][u])
__constant__ float2 c_cosq2[32];    // 0.25 * (cos[y][2vp], cos[y][2vp+1])
__constant__ float c_q[64];         // quantisation step * C(u)C(v), natural order

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t pack_sat_s8(int a, int b, uint32_t c)
{
    uint32_t d;
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// ---- one block per thread (the product's layout, natural-order input) ------------------------------
__global__ void __launch_bounds__(128, 4) k_thread_per_block(const int16_t* __restrict__ coef, uint8_t* __restrict__ out, uint32_t n)
{
    const uint32_t b = blockIdx.x * 128 + threadIdx.x;
    if (b >= n) return;
    const uint4* cp = (const uint4*)(coef + (size_t)b * 64);
    uint4 c[8];
#pragma unroll
    for (int i = 0; i < 8; i++) c[i] = cp[i];
    const int16_t* cs = (const int16_t*)c;
    float2 bp2[32];
#pragma unroll
    for (int v = 0; v < 8; v++)
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const float f = __fmul_rn((float)cs[8 * v + u], c_q[8 * v + u]);
            if (v & 1) bp2[(v >> 1) * 8 + u].y = f; else bp2[(v >> 1) * 8 + u].x = f;
        }
    uint32_t row_lo[8], row_hi[8];
#pragma unroll
    for (int y = 0; y < 8; y++) { row_lo[y] = 0; row_hi[y] = 0; }
#pragma unroll
    for (int xq = 0; xq < 4; xq++) {
        const int xp = xq ^ 1;
        int iv[8][2];
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int x = 2 * xp + e;
            float2 r[4];
#pragma unroll
            for (int vp = 0; vp < 4; vp++) {
                float2 acc = bp2[vp * 8];
#pragma unroll
                for (int u = 1; u < 8; u++) acc = __ffma2_rn(bp2[vp * 8 + u], c_cos2[x * 8 + u], acc);
                r[vp] = acc;
            }
#pragma unroll
            for (int y = 0; y < 8; y++) {
                float2 a2 = __fmul2_rn(r[0], c_cosq2[y * 4]);
#pragma unroll
                for (int vp = 1; vp < 4; vp++) a2 = __ffma2_rn(r[vp], c_cosq2[y * 4 + vp], a2);
                iv[y][e] = __float2int_rz(a2.x + a2.y);
            }
        }
#pragma unroll
        for (int y = 0; y < 8; y++) {
            if (xp < 2) row_lo[y] = pack_sat_s8(iv[y][1], iv[y][0], row_lo[y]);
            else        row_hi[y] = pack_sat_s8(iv[y][1], iv[y][0], row_hi[y]);
        }
    }
    uint2* o = (uint2*)(out + (size_t)b * 64);
#pragma unroll
    for (int y = 0; y < 8; y++) o[y] = make_uint2(row_lo[y] ^ 0x80808080u, row_hi[y] ^ 0x80808080u);
}

// ---- four lanes per block: lane q owns coefficient rows 2q, 2q+1, then output columns 2q, 2q+1 ----------
__global__ void __launch_bounds__(128) k_four_lanes_per_block(const int16_t* __restrict__ coef, uint8_t* __restrict__ out, uint32_t n)
{
    __shared__ float2 s_r[32 * 36];                        // [block of the CTA][row pair][x], padded (36 / 9) against bank conflicts
    const uint32_t lane4 = threadIdx.x & 3, lb = threadIdx.x >> 2;
    const uint32_t b = blockIdx.x * 32 + lb;
    const bool live = b < n;
    uint4 c[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
    if (live) {
        const uint4* cp = (const uint4*)(coef + (size_t)b * 64 + 16 * lane4);
        c[0] = cp[0]; c[1] = cp[1];
    }
    const int16_t* cs = (const int16_t*)c;
    float2 bp2[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
        // the two rows of this lane: natural indices 16q + u and 16q + 8 + u (the step table is indexed at run time)
        bp2[u].x = __fmul_rn((float)cs[u], c_q[16 * lane4 + u]);
        bp2[u].y = __fmul_rn((float)cs[8 + u], c_q[16 * lane4 + 8 + u]);
    }
#pragma unroll
    for (int x = 0; x < 8; x++) {
        float2 acc = bp2[0];
#pragma unroll
        for (int u = 1; u < 8; u++) acc = __ffma2_rn(bp2[u], c_cos2[x * 8 + u], acc);
        s_r[lb * 36 + lane4 * 9 + x] = acc;
    }
    __syncwarp();
    uint32_t colw[8];                                       // per row: the two bytes of this lane's columns
#pragma unroll
    for (int y = 0; y < 8; y++) colw[y] = 0;
#pragma unroll
    for (int e = 0; e < 2; e++) {
        float2 r[4];
#pragma unroll
        for (int vp = 0; vp < 4; vp++) r[vp] = s_r[lb * 36 + vp * 9 + 2 * lane4 + e];
#pragma unroll
        for (int y = 0; y < 8; y++) {
            float2 a2 = __fmul2_rn(r[0], c_cosq2[y * 4]);
#pragma unroll
            for (int vp = 1; vp < 4; vp++) a2 = __ffma2_rn(r[vp], c_cosq2[y * 4 + vp], a2);
            const int v = min(max(__float2int_rz(a2.x + a2.y) + 128, 0), 255);
            colw[y] |= (uint32_t)v << (8 * e);
        }
    }
    if (live) {
        uint16_t* o = (uint16_t*)(out + (size_t)b * 64) + lane4;
#pragma unroll
        for (int y = 0; y < 8; y++) o[4 * y] = (uint16_t)colw[y];
    }
}

int main(int argc, char** argv)
{
    const uint32_t n = argc > 1 ? (uint32_t)atol(argv[1]) : 8388608u;
    float cosv[64], q[64];
    for (int p = 0; p < 8; p++)
        for (int k = 0; k < 8; k++) cosv[p * 8 + k] = cosf(((2 * p + 1) * k * 3.14f) / 16);
    float2 cos2[64], cosq2[32];
    for (int i = 0; i < 64; i++) cos2[i] = make_float2(cosv[i], cosv[i]);
    for (int y = 0; y < 8; y++)
        for (int vp = 0; vp < 4; vp++) cosq2[y * 4 + vp] = make_float2(0.25f * cosv[y * 8 + 2 * vp], 0.25f * cosv[y * 8 + 2 * vp + 1]);
    for (int v = 0; v < 8; v++)
        for (int u = 0; u < 8; u++) q[8 * v + u] = (float)(3 + u + v) * ((u == 0 ? 0.70710677f : 1.f) * (v == 0 ? 0.70710677f : 1.f));
    CK(cudaMemcpyToSymbol(c_cos2, cos2, sizeof cos2));
    CK(cudaMemcpyToSymbol(c_cosq2, cosq2, sizeof cosq2));
    CK(cudaMemcpyToSymbol(c_q, q, sizeof q));

    std::vector<int16_t> h((size_t)1 << 22);                // 4 Mi coefficients of pattern, tiled over the slab
    uint32_t s = 12345;
    for (size_t i = 0; i < h.size(); i++) {
        s = s * 1664525u + 1013904223u;
        const int pos = (int)(i & 63), u = pos & 7, v = pos >> 3;
        const bool nz = (s >> 8) % 100 < (uint32_t)(pos == 0 ? 100 : 60 / (1 + u + v));
        h[i] = nz ? (int16_t)((int)((s >> 16) % 61) - 30) * (pos == 0 ? 8 : 1) : 0;
    }
    int16_t* d_coef; uint8_t *d_a, *d_b;
    CK(cudaMalloc(&d_coef, (size_t)n * 128));
    CK(cudaMalloc(&d_a, (size_t)n * 64));
    CK(cudaMalloc(&d_b, (size_t)n * 64));
    for (size_t off = 0; off < (size_t)n * 64; off += h.size())
        CK(cudaMemcpy(d_coef + off, h.data(), sizeof(int16_t) * (off + h.size() <= (size_t)n * 64 ? h.size() : (size_t)n * 64 - off), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms[2] = {0, 0};
    for (int which = 0; which < 2; which++) {
        for (int rep = 0; rep < 8; rep++) {
            if (rep == 3) CK(cudaEventRecord(e0));
            if (which == 0) k_thread_per_block<<<(n + 127) / 128, 128>>>(d_coef, d_a, n);
            else k_four_lanes_per_block<<<(n + 31) / 32, 128>>>(d_coef, d_b, n);
        }
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms[which], e0, e1));
        ms[which] /= 5;
    }
    std::vector<uint8_t> a((size_t)1 << 24), b2((size_t)1 << 24);
    CK(cudaMemcpy(a.data(), d_a, a.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b2.data(), d_b, b2.size(), cudaMemcpyDeviceToHost));
    size_t diff = 0;
    for (size_t i = 0; i < a.size(); i++) diff += a[i] != b2[i];
    int regs[2] = {0, 0}, occ[2] = {0, 0};
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, k_thread_per_block)); regs[0] = fa.numRegs;
    CK(cudaFuncGetAttributes(&fa, k_four_lanes_per_block)); regs[1] = fa.numRegs;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[0], k_thread_per_block, 128, 0));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[1], k_four_lanes_per_block, 128, 0));
    printf("{\"blocks\": %u, \"thread_per_block\": {\"ms\": %.4f, \"registers\": %d, \"warps_per_sm\": %d, \"Gblocks_per_s\": %.2f}, "
           "\"four_lanes_per_block\": {\"ms\": %.4f, \"registers\": %d, \"warps_per_sm\": %d, \"Gblocks_per_s\": %.2f}, "
           "\"ratio_4lane_over_thread\": %.3f, \"outputs_differ_in_first_16MB\": %zu}\n",
           n, ms[0], regs[0], occ[0] * 4, n / ms[0] / 1e6, ms[1], regs[1], occ[1] * 4, n / ms[1] / 1e6, ms[1] / ms[0], diff);
    return 0;
}
