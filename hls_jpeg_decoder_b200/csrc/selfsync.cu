// csrc/selfsync.cu -- kernel 1b: entropy decode INSIDE a restart-free scan.
//
// The reference decodes one bitstream strictly serially through a single bit reservoir
// (ProcessHuffmanBlock, loadjpg.cpp:497-863, driven block after block by DecodeMCU 945-997).
// Without restart markers nothing in the stream says where a symbol starts, so this path
// speculates and lets Huffman codes' self-synchronisation do the rest:
//
//   0. de-stuff: FF00 -> FF once, in parallel (count / prefix-sum / scatter), so that bit positions
//      are plain offsets and the decoders carry no marker logic;
//   1. every thread decodes one 128-byte sub-sequence starting at its fixed bit offset, assuming a
//      block starts there, and records its exit state (bit position, zig-zag index, block-in-MCU);
//   2. sync rounds: a thread re-decodes its sub-sequence from its left neighbour's exit state whenever
//      that state differs from the one it last used.  Inside a warp the neighbour state travels by
//      warp shuffle and the round iterates until the warp is stable; across warps it travels through
//      HBM and the host repeats the round until no exit state moves.  Sub-sequence 0 starts from the
//      true state, so by induction the fixed point is exactly the sequential decode;
//   3. an exclusive prefix sum of the per-sub-sequence block counts gives every thread its first
//      output block;
//   4. the final pass decodes from the correct entry states and writes int16 coefficients (DC still
//      as differences) into the zero-filled slab;
//   5. a per-component prefix sum over MCUs turns DC differences into DC values
//      (DCT[0] = data + prevDC in int16, loadjpg.cpp:664-665).
// Output: the same dense [block][64] zig-zag int16 layout kernel 1a produces.
#include "selfsync.cuh"
#include "device_common.cuh"
#include <cuda_runtime.h>

// ------------------------------------------------------------------------------------------
// device-wide exclusive scan (uint32, wrap-around arithmetic)
// ------------------------------------------------------------------------------------------
#define SCAN_THREADS 256
#define SCAN_ITEMS   8
#define SCAN_TILE    (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t* total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t before = 0, sum = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); k++) {
        const uint32_t c = s_warp[k];
        if (k < warp) before += c;
        sum += c;
    }
    __syncthreads();
    *total = sum;
    return before + incl - v;
}

__global__ void __launch_bounds__(SCAN_THREADS)
hjd_k_scan_tiles(uint32_t* __restrict__ data, uint32_t n, uint32_t* __restrict__ tile_sums)
{
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) { v[i] = (base + i < n) ? data[base + i] : 0u; sum += v[i]; }
    uint32_t total;
    uint32_t run = block_exclusive_scan(sum, s_warp, &total);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) { if (base + i < n) data[base + i] = run; run += v[i]; }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024)
hjd_k_scan_sums(uint32_t* __restrict__ tile_sums, uint32_t n_tiles)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_tiles; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = (i < n_tiles) ? tile_sums[i] : 0u;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, s_warp, &total);
        const uint32_t carry = s_carry;
        if (i < n_tiles) tile_sums[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SCAN_THREADS)
hjd_k_scan_add(uint32_t* __restrict__ data, uint32_t n, const uint32_t* __restrict__ tile_sums)
{
    const uint32_t add = tile_sums[blockIdx.x];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++)
        if (base + i < n) data[base + i] += add;
}

cudaError_t hjd_scan_u32(uint32_t* data, uint32_t n, uint32_t* tmp, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    const uint32_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    hjd_k_scan_tiles<<<tiles, SCAN_THREADS, 0, st>>>(data, n, tmp);
    if (tiles > 1) {
        hjd_k_scan_sums<<<1, 1024, 0, st>>>(tmp, tiles);
        hjd_k_scan_add<<<tiles, SCAN_THREADS, 0, st>>>(data, n, tmp);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------
// Index of the HjdSsImage whose [base, base + count) range (selected by FIELD) contains key.
template <int FIELD>   // 0: sub-sequences, 1: chunks, 2: MCUs
__device__ __forceinline__ int ss_find(const HjdSsImage* __restrict__ ss, int n_ss, uint32_t key)
{
    int lo = 0, hi = n_ss - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        const uint32_t b = FIELD == 0 ? ss[mid].sub_base : (FIELD == 1 ? ss[mid].chunk_base : ss[mid].mcu_base);
        if (b <= key) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// ------------------------------------------------------------------------------------------
// step 0: de-stuffing
// ------------------------------------------------------------------------------------------
// A thread owns one 16-byte aligned chunk of the stuffed scan.  A byte is dropped iff it is 00 and
// its predecessor (inside the scan) is FF.  Returns the keep mask (bit j = keep byte j).
__device__ __forceinline__ uint32_t destuff_mask(const uint8_t* a0, uint32_t lc, uint32_t lead, uint32_t scan_len,
                                                 uint4* out_bytes)
{
    const uint4 v = __ldg((const uint4*)(a0 + (size_t)lc * 16));
    *out_bytes = v;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    const uint32_t q0 = lc * 16;                       // position of byte 0 relative to a0
    uint32_t prev = (q0 > lead) ? a0[(size_t)q0 - 1] : 0u;
    uint32_t keep = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const uint32_t b = (w[j >> 2] >> (8 * (j & 3))) & 255u;
        const uint32_t q = q0 + j;
        const bool valid = (q >= lead) && (q - lead < scan_len);
        const bool stuffed = (b == 0u) && (prev == 0xFFu) && (q > lead);
        if (valid && !stuffed) keep |= 1u << j;
        prev = b;
    }
    return keep;
}

__global__ void __launch_bounds__(256)
hjd_k_destuff_count(const uint8_t* __restrict__ arena, const HjdImageDesc* __restrict__ imgs,
                    const HjdSsImage* __restrict__ ss, int n_ss, uint32_t n_chunks_total,
                    uint32_t* __restrict__ counts)
{
    const uint32_t t = blockIdx.x * 256 + threadIdx.x;
    if (t >= n_chunks_total) { if (t == n_chunks_total) counts[t] = 0; return; }
    const HjdSsImage s = ss[ss_find<1>(ss, n_ss, t)];
    const HjdImageDesc* d = imgs + s.img;
    const uint8_t* a0 = arena + d->scan_off - s.lead;
    uint4 bytes;
    counts[t] = __popc(destuff_mask(a0, t - s.chunk_base, s.lead, d->scan_len, &bytes));
}

__global__ void __launch_bounds__(256)
hjd_k_destuff_scatter(const uint8_t* __restrict__ arena, const HjdImageDesc* __restrict__ imgs,
                      const HjdSsImage* __restrict__ ss, int n_ss, uint32_t n_chunks_total,
                      const uint32_t* __restrict__ prefix, uint8_t* __restrict__ dst, uint32_t* __restrict__ dlen)
{
    const uint32_t t = blockIdx.x * 256 + threadIdx.x;
    if (t >= n_chunks_total) return;
    const int si = ss_find<1>(ss, n_ss, t);
    const HjdSsImage s = ss[si];
    const HjdImageDesc* d = imgs + s.img;
    const uint8_t* a0 = arena + d->scan_off - s.lead;
    uint4 bytes;
    uint32_t keep = destuff_mask(a0, t - s.chunk_base, s.lead, d->scan_len, &bytes);
    const uint32_t w[4] = {bytes.x, bytes.y, bytes.z, bytes.w};
    uint8_t* out = dst + s.dst_off + (prefix[t] - prefix[s.chunk_base]);
#pragma unroll
    for (int j = 0; j < 16; j++)
        if (keep & (1u << j)) *out++ = (uint8_t)(w[j >> 2] >> (8 * (j & 3)));
    if (t == s.chunk_base + s.n_chunks - 1) {          // last chunk of the image: length + zero slack
        const uint32_t len = prefix[t] - prefix[s.chunk_base] + __popc(keep);
        dlen[si] = len;
        uint8_t* z = dst + s.dst_off + len;
        for (int j = 0; j < HJD_SS_SLACK; j++) z[j] = 0;
    }
}

cudaError_t hjd_launch_destuff(const uint8_t* arena, const HjdImageDesc* imgs, const HjdSsImage* ss, int n_ss,
                               uint32_t n_chunks_total, uint32_t* counts, uint32_t* scan_tmp,
                               uint8_t* dst, uint32_t* dlen, cudaStream_t st)
{
    if (n_ss <= 0 || n_chunks_total == 0) return cudaSuccess;
    const uint32_t grid = (n_chunks_total + 1 + 255) / 256;
    hjd_k_destuff_count<<<grid, 256, 0, st>>>(arena, imgs, ss, n_ss, n_chunks_total, counts);
    cudaError_t e = hjd_scan_u32(counts, n_chunks_total + 1, scan_tmp, st);
    if (e != cudaSuccess) return e;
    hjd_k_destuff_scatter<<<grid, 256, 0, st>>>(arena, imgs, ss, n_ss, n_chunks_total, counts, dst, dlen);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// the sub-sequence decoder shared by the sync rounds and the final pass
// ------------------------------------------------------------------------------------------
// state = bit position (40 bits) | zig-zag index (7 bits) << 40 | block-in-MCU index (4 bits) << 47
#define SS_INVALID 0xFFFFFFFFFFFFFFFFull
__device__ __forceinline__ uint64_t ss_pack(uint64_t p, int k, int c) { return p | ((uint64_t)k << 40) | ((uint64_t)c << 47); }

struct SsCtx {
    const uint8_t* D;       // de-stuffed stream (zero slack after its end)
    uint32_t sh_tab;        // shared-window address of the component tables (DC, AC per component)
    uint32_t bpm, ny;
};

__device__ __forceinline__ uint32_t ss_load_be32(const uint8_t* D, uint32_t bpos)
{
    const uintptr_t a = (uintptr_t)(D + bpos);
    const uint32_t* ap = (const uint32_t*)(a & ~(uintptr_t)3);
    const uint32_t x = __funnelshift_r(__ldg(ap), __ldg(ap + 1), (uint32_t)(a & 3) * 8);
    return __byte_perm(x, 0, 0x0123);
}

// Decode symbols from `state` until the bit position reaches end_bit (sync rounds) or, when WRITE,
// until the image's last block is complete.  Same symbol semantics as kernel 1a / the reference
// (loadjpg.cpp:559-829); an undecodable code consumes one bit so that every path makes progress.
template <bool WRITE>
__device__ __forceinline__ uint64_t ss_decode(const SsCtx& cx, uint64_t state, uint64_t end_bit, uint32_t* nb_out,
                                              int16_t* coef_img, uint32_t blk0, uint32_t n_blocks, int* flags)
{
    constexpr uint32_t kTabBytes = (uint32_t)sizeof(HjdHuffTable);
    uint64_t p = state & 0xFFFFFFFFFFull;
    int k = (int)((state >> 40) & 127u), c = (int)((state >> 47) & 15u);
    uint32_t bpos = (uint32_t)(p >> 3);
    uint32_t hi = ss_load_be32(cx.D, bpos), lo = ss_load_be32(cx.D, bpos + 4);
    bpos += 8;
    const uint32_t sh0 = (uint32_t)p & 7u;
    hi = __funnelshift_l(lo, hi, sh0);
    lo <<= sh0;
    int nbits = 64 - (int)sh0;
    uint32_t nb = 0;
    uint32_t tbase = cx.sh_tab + ((uint32_t)c < cx.ny ? 0u : ((uint32_t)c == cx.ny ? 2u : 4u)) * kTabBytes;

    while (p < end_bit && (!WRITE || blk0 + nb < n_blocks)) {
        if (nbits < 32) {
            const uint32_t w = ss_load_be32(cx.D, bpos);
            bpos += 4;
            hi |= hjd_shr(w, (uint32_t)nbits);                 // nbits in [1, 31]
            lo |= hjd_shl(w, 32u - (uint32_t)nbits);
            nbits += 32;
        }
        const uint32_t is_ac = (uint32_t)min(k, 1);
        const uint32_t t = tbase + is_ac * kTabBytes;
        uint32_t e = hjd_lds_u16(t + ((hi >> (32 - HJD_LUT_BITS)) << 1));
        if ((e & 31u) == 0) {
            e = hjd_long_code(t, hi >> 16, is_ac != 0);
            if (e == 0) { e = hjd_sym_fields(1, 0, is_ac != 0); *flags |= HJD_ST_BAD_CODE; }   // consume one bit
        }
        const uint32_t len = e & 31u, size = (e >> 5) & 15u, kadv = (e >> 9) & 127u;
        const bool store = size != 0u;              // a size-0 DC difference adds nothing; the slab is pre-zeroed
        const uint32_t after = __funnelshift_l(lo, hi, len);
        const uint32_t v = hjd_shr(after, 32u - size);
        const int neg = ~((int)after >> 31);
        const int val = (int)v + (neg & (int)(hjd_shl(0xFFFFFFFFu, size) + 1u));     // DetermineSign, loadjpg.cpp:396-409
        const uint32_t used = len + size;
        hi = __funnelshift_l(lo, hi, used);
        lo <<= used;
        nbits -= (int)used;
        p += used;
        const uint32_t kpos = (uint32_t)k + kadv - 1u;
        if (WRITE && store) {
            if (kpos <= 63u) coef_img[(size_t)(blk0 + nb) * 64 + kpos] = (int16_t)val;   // DC: the difference
            else *flags |= HJD_ST_COEF_RANGE;
        }
        k += (int)kadv;
        if (k >= 64) {
            nb++;
            k = 0;
            c = (c + 1 == (int)cx.bpm) ? 0 : c + 1;
            tbase = cx.sh_tab + ((uint32_t)c < cx.ny ? 0u : ((uint32_t)c == cx.ny ? 2u : 4u)) * kTabBytes;
        }
    }
    *nb_out = nb;
    return ss_pack(p, k, c);
}

// Loads the three per-component (DC, AC) table pairs of a table set into shared memory.
__device__ __forceinline__ void ss_load_tables(const HjdTableSet* ts, uint8_t* s_tab)
{
    constexpr int n16 = (int)(sizeof(HjdHuffTable) / 16);
    for (int c = 0; c < 3; c++) {
        const uint4* sdc = (const uint4*)&ts->tab[ts->dc_of_comp[c]];
        const uint4* sac = (const uint4*)&ts->tab[ts->ac_of_comp[c]];
        uint4* ddc = (uint4*)(s_tab + (2 * c) * sizeof(HjdHuffTable));
        uint4* dac = (uint4*)(s_tab + (2 * c + 1) * sizeof(HjdHuffTable));
        for (int i = threadIdx.x; i < n16; i += blockDim.x) { ddc[i] = __ldg(sdc + i); dac[i] = __ldg(sac + i); }
    }
}

// ------------------------------------------------------------------------------------------
// steps 1-2: speculative decode + synchronisation rounds
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HJD_SS_THREADS)
hjd_k_ss_round(const HjdImageDesc* __restrict__ imgs, const HjdTableSet* __restrict__ tsets,
               const HjdSsImage* __restrict__ ss, const HjdSsWork* __restrict__ work,
               const uint8_t* __restrict__ dst, const uint32_t* __restrict__ dlen, int first,
               const uint64_t* __restrict__ e_in, uint64_t* __restrict__ e_out, uint64_t* __restrict__ x_arr,
               uint32_t* __restrict__ nb_arr, int* __restrict__ changed)
{
    extern __shared__ __align__(16) uint8_t s_tab[];
    const HjdSsWork wk = work[blockIdx.x];
    const HjdSsImage s = ss[wk.ss];
    const HjdImageDesc* d = imgs + s.img;
    ss_load_tables(tsets + d->table_set, s_tab);
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const uint32_t li = wk.first_sub + threadIdx.x;                  // local sub-sequence index
    const uint32_t L = dlen[wk.ss];
    const bool active = li < s.n_subs && (uint64_t)li * HJD_SS_SUB_BYTES < L;
    const uint32_t gi = s.sub_base + li;
    SsCtx cx;
    cx.D = dst + s.dst_off;
    cx.sh_tab = (uint32_t)__cvta_generic_to_shared(s_tab);
    cx.bpm = d->blocks_per_mcu;
    cx.ny = d->ncomp == 3 ? (uint32_t)d->hf * d->vf : 1u;
    const uint64_t end_bit = (uint64_t)(li + 1) * HJD_SS_SUB_BYTES * 8;

    uint64_t e = 0, x_used = SS_INVALID;
    uint32_t nb = 0;
    if (active && !first) { e = e_in[gi]; x_used = x_arr[gi]; nb = nb_arr[gi]; }
    const uint64_t e_start = e;
    // entry state that does not change during this launch: the true start, or (lane 0) the previous
    // warp's exit state of the previous round, or the speculative guess of the very first pass
    uint64_t x_fixed = ss_pack((uint64_t)li * HJD_SS_SUB_BYTES * 8, 0, 0);     // guess: a block starts here
    if (li == 0) x_fixed = ss_pack(0, 0, 0);
    else if (!first && active && lane == 0) x_fixed = e_in[gi - 1];
    int flags = 0;

    for (int iter = 0; iter < 33; iter++) {
        const uint64_t from_left = __shfl_up_sync(0xffffffffu, e, 1);
        uint64_t xin = x_fixed;
        if (li != 0 && lane != 0 && !(first && iter == 0)) xin = from_left;
        const bool need = active && xin != x_used;
        if (need) {
            e = ss_decode<false>(cx, xin, end_bit, &nb, nullptr, 0, 0, &flags);
            x_used = xin;
        }
        if (!__any_sync(0xffffffffu, need)) break;
    }
    if (active) {
        e_out[gi] = e;
        x_arr[gi] = x_used;
        nb_arr[gi] = nb;
        if (first || e != e_start) *changed = 1;
    } else if (li < s.n_subs) {
        e_out[gi] = 0; x_arr[gi] = SS_INVALID; nb_arr[gi] = 0;
    }
}

cudaError_t hjd_launch_ss_round(const HjdImageDesc* imgs, const HjdTableSet* tsets, const HjdSsImage* ss,
                                const HjdSsWork* work, int n_work, const uint8_t* dst, const uint32_t* dlen,
                                int first, const uint64_t* e_in, uint64_t* e_out, uint64_t* x, uint32_t* nb,
                                int* changed, cudaStream_t st)
{
    if (n_work <= 0) return cudaSuccess;
    const size_t smem = 6 * sizeof(HjdHuffTable);
    hjd_k_ss_round<<<n_work, HJD_SS_THREADS, smem, st>>>(imgs, tsets, ss, work, dst, dlen, first, e_in, e_out, x, nb, changed);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// step 4: final pass with output
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HJD_SS_THREADS)
hjd_k_ss_write(const HjdImageDesc* __restrict__ imgs, const HjdTableSet* __restrict__ tsets,
               const HjdSsImage* __restrict__ ss, const HjdSsWork* __restrict__ work,
               const uint8_t* __restrict__ dst, const uint32_t* __restrict__ dlen,
               const uint64_t* __restrict__ x_arr, const uint32_t* __restrict__ first_block,
               int16_t* __restrict__ coef, int32_t* __restrict__ status)
{
    extern __shared__ __align__(16) uint8_t s_tab[];
    const HjdSsWork wk = work[blockIdx.x];
    const HjdSsImage s = ss[wk.ss];
    const HjdImageDesc* d = imgs + s.img;
    ss_load_tables(tsets + d->table_set, s_tab);
    __syncthreads();

    const uint32_t li = wk.first_sub + threadIdx.x;
    const uint32_t L = dlen[wk.ss];
    if (li >= s.n_subs || (uint64_t)li * HJD_SS_SUB_BYTES >= L) return;
    const uint32_t gi = s.sub_base + li;
    SsCtx cx;
    cx.D = dst + s.dst_off;
    cx.sh_tab = (uint32_t)__cvta_generic_to_shared(s_tab);
    cx.bpm = d->blocks_per_mcu;
    cx.ny = d->ncomp == 3 ? (uint32_t)d->hf * d->vf : 1u;
    const uint64_t end_bit = (uint64_t)(li + 1) * HJD_SS_SUB_BYTES * 8;
    const uint32_t blk0 = first_block[gi] - first_block[s.sub_base];
    const uint32_t n_blocks = (uint32_t)d->n_blocks;
    int flags = 0;
    uint32_t nb = 0;
    const uint64_t xs = x_arr[gi];
    if (xs == SS_INVALID) flags |= HJD_ST_BAD_CODE;                 // the rounds did not reach this sub-sequence
    else ss_decode<true>(cx, xs, end_bit, &nb, coef + d->block_base * 64, blk0, n_blocks, &flags);
    // the last sub-sequence that holds data must complete the image
    const bool last = (li + 1 == s.n_subs) || ((uint64_t)(li + 1) * HJD_SS_SUB_BYTES >= L);
    if (last && blk0 + nb < n_blocks) flags |= HJD_ST_OVERRUN;
    if (flags) atomicOr(&status[s.img], flags);
}

cudaError_t hjd_launch_ss_write(const HjdImageDesc* imgs, const HjdTableSet* tsets, const HjdSsImage* ss,
                                const HjdSsWork* work, int n_work, const uint8_t* dst, const uint32_t* dlen,
                                const uint64_t* x, const uint32_t* first_block, int16_t* coef, int32_t* status,
                                cudaStream_t st)
{
    if (n_work <= 0) return cudaSuccess;
    const size_t smem = 6 * sizeof(HjdHuffTable);
    hjd_k_ss_write<<<n_work, HJD_SS_THREADS, smem, st>>>(imgs, tsets, ss, work, dst, dlen, x, first_block, coef, status);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// step 5: DC differences -> DC values
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
hjd_k_dc_sums(const HjdImageDesc* __restrict__ imgs, const HjdSsImage* __restrict__ ss, int n_ss,
              uint32_t n_mcus_total, const int16_t* __restrict__ coef, uint32_t* __restrict__ sums)
{
    const uint32_t m = blockIdx.x * 256 + threadIdx.x;
    if (m >= n_mcus_total) return;
    const HjdSsImage s = ss[ss_find<2>(ss, n_ss, m)];
    const HjdImageDesc* d = imgs + s.img;
    const uint32_t bpm = d->blocks_per_mcu, ny = d->ncomp == 3 ? (uint32_t)d->hf * d->vf : 1u;
    const int16_t* c0 = coef + (d->block_base + (uint64_t)(m - s.mcu_base) * bpm) * 64;
    uint32_t sy = 0;
    for (uint32_t j = 0; j < ny; j++) sy += (uint32_t)(int)c0[(size_t)j * 64];
    sums[m] = sy;
    sums[n_mcus_total + m] = bpm > 1 ? (uint32_t)(int)c0[(size_t)ny * 64] : 0u;
    sums[2 * n_mcus_total + m] = bpm > 1 ? (uint32_t)(int)c0[(size_t)(ny + 1) * 64] : 0u;
}

__global__ void __launch_bounds__(256)
hjd_k_dc_apply(const HjdImageDesc* __restrict__ imgs, const HjdSsImage* __restrict__ ss, int n_ss,
               uint32_t n_mcus_total, const uint32_t* __restrict__ prefix, int16_t* __restrict__ coef)
{
    const uint32_t m = blockIdx.x * 256 + threadIdx.x;
    if (m >= n_mcus_total) return;
    const HjdSsImage s = ss[ss_find<2>(ss, n_ss, m)];
    const HjdImageDesc* d = imgs + s.img;
    const uint32_t bpm = d->blocks_per_mcu, ny = d->ncomp == 3 ? (uint32_t)d->hf * d->vf : 1u;
    int16_t* c0 = coef + (d->block_base + (uint64_t)(m - s.mcu_base) * bpm) * 64;
    uint32_t run = prefix[m] - prefix[s.mcu_base];                        // predictor before this MCU (mod 2^16 matters)
    for (uint32_t j = 0; j < ny; j++) {
        run += (uint32_t)(int)c0[(size_t)j * 64];
        c0[(size_t)j * 64] = (int16_t)run;                                // DCT[0] = data + prevDC, loadjpg.cpp:664
    }
    if (bpm > 1) {
        const uint32_t pb = prefix[n_mcus_total + m] - prefix[n_mcus_total + s.mcu_base];
        const uint32_t pr = prefix[2 * n_mcus_total + m] - prefix[2 * n_mcus_total + s.mcu_base];
        c0[(size_t)ny * 64] = (int16_t)(pb + (uint32_t)(int)c0[(size_t)ny * 64]);
        c0[(size_t)(ny + 1) * 64] = (int16_t)(pr + (uint32_t)(int)c0[(size_t)(ny + 1) * 64]);
    }
}

cudaError_t hjd_launch_dc_sums(const HjdImageDesc* imgs, const HjdSsImage* ss, int n_ss, uint32_t n_mcus_total,
                               const int16_t* coef, uint32_t* sums, cudaStream_t st)
{
    if (n_ss <= 0 || n_mcus_total == 0) return cudaSuccess;
    hjd_k_dc_sums<<<(n_mcus_total + 255) / 256, 256, 0, st>>>(imgs, ss, n_ss, n_mcus_total, coef, sums);
    return cudaGetLastError();
}

cudaError_t hjd_launch_dc_apply(const HjdImageDesc* imgs, const HjdSsImage* ss, int n_ss, uint32_t n_mcus_total,
                                const uint32_t* prefix, int16_t* coef, cudaStream_t st)
{
    if (n_ss <= 0 || n_mcus_total == 0) return cudaSuccess;
    hjd_k_dc_apply<<<(n_mcus_total + 255) / 256, 256, 0, st>>>(imgs, ss, n_ss, n_mcus_total, prefix, coef);
    return cudaGetLastError();
}
