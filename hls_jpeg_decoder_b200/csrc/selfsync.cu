// csrc/selfsync.cu -- kernel 1b: entropy decode INSIDE a restart-free scan.
//
// The reference decodes one bitstream strictly serially through a single bit reservoir
// (ProcessHuffmanBlock, loadjpg.cpp:497-863, driven block after block by DecodeMCU 945-997).
// Without restart markers nothing in the stream says where a symbol starts, so this path
// speculates and lets Huffman codes' self-synchronisation do the rest:
//
//   0. de-stuff: FF00 -> FF once, in parallel (count / prefix-sum / scatter), so that bit positions
//      are plain offsets and the decoders carry no marker logic;
//   1. every thread decodes from a fixed bit offset one sub-sequence ahead of its own 128-byte
//      sub-sequence (assuming a block starts there), then its own, and records its exit state
//      (bit position, zig-zag index, block-in-MCU);
//   2. sync rounds: a thread re-decodes its sub-sequence from its left neighbour's exit state whenever
//      that state differs from the one it last used.  Inside a warp the neighbour state travels by
//      warp shuffle and the round iterates until the warp is stable; across warps it travels through
//      HBM and the host repeats the round until no exit state moves.  Sub-sequence 0 starts from the
//      true state, so by induction the fixed point is exactly the sequential decode;
//   3. an exclusive prefix sum of, per sub-sequence, the number of blocks that start in it and the sums
//      of their DC differences gives every thread its first output block and its DC predictors
//      (DCT[0] = data + prevDC in int16, loadjpg.cpp:664-665);
//   4. the final pass decodes, from the correct entry states, the blocks that start in each
//      sub-sequence and writes every block once, in full, as one 128-byte line.
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
// the sub-sequence decoder of the synchronisation rounds
// ------------------------------------------------------------------------------------------
// state = bit position (40 bits) | zig-zag index (7 bits) << 40 | block-in-MCU index (4 bits) << 47
#define SS_INVALID 0xFFFFFFFFFFFFFFFFull
__device__ __forceinline__ uint64_t ss_pack(uint64_t p, int k, int c) { return p | ((uint64_t)k << 40) | ((uint64_t)c << 47); }

struct SsCtx {
    const uint8_t* D;       // de-stuffed stream (zero slack after its end)
    uint32_t sh_tab;        // shared-window address of the component tables (DC, AC per component)
    uint32_t bpm, ny;
};

// What a sub-sequence contributes to the output layout: blocks that START inside it (a block belongs
// to the sub-sequence in which its DC symbol begins) and the sums of their DC differences per component.
struct SsCount { uint32_t ns, dc0, dc1, dc2; };

__device__ __forceinline__ uint32_t ss_load_be32(const uint8_t* D, uint32_t bpos)
{
    const uintptr_t a = (uintptr_t)(D + bpos);
    const uint32_t* ap = (const uint32_t*)(a & ~(uintptr_t)3);
    const uint32_t x = __funnelshift_r(__ldg(ap), __ldg(ap + 1), (uint32_t)(a & 3) * 8);
    return __byte_perm(x, 0, 0x0123);
}

// One Huffman symbol from the window (hi:lo): the fields of hjd_sym_fields plus the extended value.
// An undecodable code consumes one bit (as a size-0 symbol) so that every path makes progress; the
// synchronisation rounds and the write pass must agree on this rule, which is why they share this code.
struct SsSym { uint32_t used, size, kadv; int val; bool bad; };

__device__ __forceinline__ SsSym ss_symbol(uint32_t t, uint32_t hi, uint32_t lo, bool is_ac)
{
    SsSym s;
    uint32_t e = hjd_lds_u16(t + ((hi >> (32 - HJD_LUT_BITS)) << 1));
    s.bad = false;
    if ((e & 31u) == 0) {
        e = hjd_long_code(t, hi >> 16, is_ac);
        if (e == 0) { e = hjd_sym_fields(1, 0, is_ac); s.bad = true; }
    }
    const uint32_t len = e & 31u;
    s.size = (e >> 5) & 15u;
    s.kadv = (e >> 9) & 127u;
    const uint32_t after = __funnelshift_l(lo, hi, len);
    const uint32_t v = hjd_shr(after, 32u - s.size);
    const int neg = ~((int)after >> 31);
    s.val = (int)v + (neg & (int)(hjd_shl(0xFFFFFFFFu, s.size) + 1u));       // DetermineSign, loadjpg.cpp:396-409
    s.used = len + s.size;
    return s;
}

// Decode symbols from `state` until the bit position reaches end_bit; returns the exit state.
// COUNT: also count the blocks that start in [entry, end_bit) and sum their DC differences.
template <bool COUNT>
__device__ __forceinline__ uint64_t ss_scan_decode(const SsCtx& cx, uint64_t state, uint64_t end_bit, SsCount* cnt)
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
    uint32_t comp = (uint32_t)c < cx.ny ? 0u : ((uint32_t)c == cx.ny ? 1u : 2u);
    uint32_t tbase = cx.sh_tab + comp * 2u * kTabBytes;
    uint32_t ns = 0, d0 = 0, d1 = 0, d2 = 0;

    while (p < end_bit) {
        if (nbits < 32) {
            const uint32_t w = ss_load_be32(cx.D, bpos);
            bpos += 4;
            hi |= hjd_shr(w, (uint32_t)nbits);
            lo |= hjd_shl(w, 32u - (uint32_t)nbits);
            nbits += 32;
        }
        const bool is_ac = k != 0;
        const SsSym s = ss_symbol(tbase + (is_ac ? kTabBytes : 0u), hi, lo, is_ac);
        hi = __funnelshift_l(lo, hi, s.used);
        lo <<= s.used;
        nbits -= (int)s.used;
        p += s.used;
        if (COUNT && !is_ac) {
            ns++;
            d0 += comp == 0u ? (uint32_t)s.val : 0u;
            d1 += comp == 1u ? (uint32_t)s.val : 0u;
            d2 += comp == 2u ? (uint32_t)s.val : 0u;
        }
        k += (int)s.kadv;
        if (k >= 64) {
            k = 0;
            c = (c + 1 == (int)cx.bpm) ? 0 : c + 1;
            comp = (uint32_t)c < cx.ny ? 0u : ((uint32_t)c == cx.ny ? 1u : 2u);
            tbase = cx.sh_tab + comp * 2u * kTabBytes;
        }
    }
    if (COUNT) { cnt->ns = ns; cnt->dc0 = d0; cnt->dc1 = d1; cnt->dc2 = d2; }
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
// Round 0 (first = 1): every thread decodes the sub-sequence BEFORE its own from the fixed bit offset
// (assuming a block starts there) and takes the state it arrives in as the entry state of its own
// sub-sequence: after 1024 bits of Huffman data the decoder has almost surely synchronised, so most
// entry states are already correct and independent of the neighbours.
// Later rounds: a thread whose left neighbour's exit state differs from the entry state it used
// re-decodes; inside a warp the neighbour state travels by warp shuffle and the round iterates until
// the warp is stable, across warps it travels through HBM and the host repeats the round until no exit
// state moves.  Sub-sequence 0 starts from the true state, so the fixed point is the sequential decode.
__global__ void __launch_bounds__(HJD_SS_THREADS)
hjd_k_ss_round(const HjdImageDesc* __restrict__ imgs, const HjdTableSet* __restrict__ tsets,
               const HjdSsImage* __restrict__ ss, const HjdSsWork* __restrict__ work,
               const uint8_t* __restrict__ dst, const uint32_t* __restrict__ dlen, int first, uint32_t n_subs_total,
               const uint64_t* __restrict__ e_in, uint64_t* __restrict__ e_out, uint64_t* __restrict__ x_arr,
               uint32_t* __restrict__ cnt_arr, int* __restrict__ changed)
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
    const uint64_t sub_bits = (uint64_t)HJD_SS_SUB_BYTES * 8;
    const uint64_t end_bit = (uint64_t)(li + 1) * sub_bits;

    uint64_t e = 0, x_used = SS_INVALID;
    SsCount cnt = {0, 0, 0, 0};
    bool moved = false;
    if (first) {
        if (active) {
            x_used = ss_pack(0, 0, 0);
            if (li != 0) x_used = ss_scan_decode<false>(cx, ss_pack((uint64_t)(li - 1) * sub_bits, 0, 0), (uint64_t)li * sub_bits, nullptr);
            e = ss_scan_decode<true>(cx, x_used, end_bit, &cnt);
            moved = true;
        }
    } else {
        if (active) {
            e = e_in[gi]; x_used = x_arr[gi];
            cnt.ns = cnt_arr[gi]; cnt.dc0 = cnt_arr[n_subs_total + gi];
            cnt.dc1 = cnt_arr[2 * n_subs_total + gi]; cnt.dc2 = cnt_arr[3 * n_subs_total + gi];
        }
        const uint64_t e_start = e;
        uint64_t x_fixed = ss_pack(0, 0, 0);                         // li == 0: the true start
        if (li != 0 && active && lane == 0) x_fixed = e_in[gi - 1];  // previous warp: last round's exit state
        for (int iter = 0; iter < 33; iter++) {
            const uint64_t from_left = __shfl_up_sync(0xffffffffu, e, 1);
            const uint64_t xin = (li != 0 && lane != 0) ? from_left : x_fixed;
            const bool need = active && xin != x_used;
            if (need) {
                e = ss_scan_decode<true>(cx, xin, end_bit, &cnt);
                x_used = xin;
            }
            if (!__any_sync(0xffffffffu, need)) break;
        }
        moved = active && e != e_start;
    }
    if (active) {
        e_out[gi] = e;
        x_arr[gi] = x_used;
        cnt_arr[gi] = cnt.ns;
        cnt_arr[n_subs_total + gi] = cnt.dc0;
        cnt_arr[2 * n_subs_total + gi] = cnt.dc1;
        cnt_arr[3 * n_subs_total + gi] = cnt.dc2;
        if (moved) *changed = 1;
    } else if (li < s.n_subs) {
        e_out[gi] = 0; x_arr[gi] = SS_INVALID;
        cnt_arr[gi] = 0; cnt_arr[n_subs_total + gi] = 0; cnt_arr[2 * n_subs_total + gi] = 0; cnt_arr[3 * n_subs_total + gi] = 0;
    }
}

cudaError_t hjd_launch_ss_round(const HjdImageDesc* imgs, const HjdTableSet* tsets, const HjdSsImage* ss,
                                const HjdSsWork* work, int n_work, const uint8_t* dst, const uint32_t* dlen,
                                int first, uint32_t n_subs_total, const uint64_t* e_in, uint64_t* e_out, uint64_t* x,
                                uint32_t* cnt, int* changed, cudaStream_t st)
{
    if (n_work <= 0) return cudaSuccess;
    const size_t smem = 6 * sizeof(HjdHuffTable);
    hjd_k_ss_round<<<n_work, HJD_SS_THREADS, smem, st>>>(imgs, tsets, ss, work, dst, dlen, first, n_subs_total,
                                                        e_in, e_out, x, cnt, changed);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// step 4: final pass with output
// ------------------------------------------------------------------------------------------
// Every thread decodes the blocks that START in its sub-sequence, each to its end (running into the
// next sub-sequence if need be), after skipping the tail of the block it was entered in the middle of
// (that block belongs to an earlier thread).  So every 8x8 block is produced by exactly one thread, in
// full, and leaves the SM as one 128-byte line through the same slot / list / four-at-a-time flush as
// kernel 1a; the coefficient slab needs no zero-fill.  prefix[] = exclusive scan of the counts of the
// last round: first owned block and, per component, the DC predictor at it
// (DCT[0] = data + prevDC in int16, loadjpg.cpp:664-665).
#define SS_WRITE_SYMS 4

__global__ void __launch_bounds__(HJD_SS_THREADS)
hjd_k_ss_write(const HjdImageDesc* __restrict__ imgs, const HjdTableSet* __restrict__ tsets,
               const HjdSsImage* __restrict__ ss, const HjdSsWork* __restrict__ work,
               const uint8_t* __restrict__ dst, const uint32_t* __restrict__ dlen, uint32_t n_subs_total,
               const uint64_t* __restrict__ x_arr, const uint32_t* __restrict__ prefix,
               int16_t* __restrict__ coef, int32_t* __restrict__ status)
{
    extern __shared__ __align__(16) uint8_t s_raw[];
    constexpr uint32_t kTabBytes = (uint32_t)sizeof(HjdHuffTable);
    constexpr uint32_t kSlotBytes = HJD_SS_THREADS * 128, kListBytes = HJD_SS_THREADS * 8;
    uint8_t* s_tab = s_raw + kSlotBytes + kListBytes;
    const HjdSsWork wk = work[blockIdx.x];
    const HjdSsImage s = ss[wk.ss];
    const HjdImageDesc* d = imgs + s.img;
    ss_load_tables(tsets + d->table_set, s_tab);
    {
        uint4* z = (uint4*)s_raw;
        for (int i = threadIdx.x; i < HJD_SS_THREADS * 8; i += HJD_SS_THREADS) z[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();

    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t sh_base = (uint32_t)__cvta_generic_to_shared(s_raw);
    const uint32_t sh_list = sh_base + kSlotBytes;
    const uint32_t sh_tab = sh_list + kListBytes;
    const uint32_t my_slot = sh_base + (uint32_t)tid * 128u;
    const uint32_t swz = (uint32_t)(lane & 7) << 4;
    const uint32_t warp_slots = sh_base + (uint32_t)(tid & ~31) * 128u;
    const uint32_t warp_list = sh_list + (uint32_t)(tid & ~31) * 8u;
    const uint32_t lt_mask = (1u << lane) - 1u;

    const uint32_t li = wk.first_sub + tid;
    const uint32_t L = dlen[wk.ss];
    const uint32_t bpm = d->blocks_per_mcu, ny = d->ncomp == 3 ? (uint32_t)d->hf * d->vf : 1u;
    const uint32_t n_blocks = (uint32_t)d->n_blocks;
    const uint32_t blk_base = (uint32_t)d->block_base;
    const uint8_t* D = dst + s.dst_off;
    const uint64_t end_bit = (uint64_t)(li + 1) * HJD_SS_SUB_BYTES * 8;

    bool finished = true;
    bool owned = false;               // false while skipping the tail of a block entered in the middle
    uint64_t p = 0;
    int k = 0, c = 0;
    uint32_t blk = 0;                 // image-local index of the block being decoded (when owned)
    int p0 = 0, p1 = 0, p2 = 0;       // DC predictors, p0 = current component
    uint32_t t0 = sh_tab, t1 = sh_tab + 2 * kTabBytes, t2 = sh_tab + 4 * kTabBytes;
    uint32_t hi = 0, lo = 0, bpos = 0;
    int nbits = 0, flags = 0;
    if (li < s.n_subs && (uint64_t)li * HJD_SS_SUB_BYTES < L) {
        const uint32_t gi = s.sub_base + li;
        const uint64_t xs = x_arr[gi];
        if (xs == SS_INVALID) flags |= HJD_ST_BAD_CODE;
        else {
            p = xs & 0xFFFFFFFFFFull;
            k = (int)((xs >> 40) & 127u);
            c = (int)((xs >> 47) & 15u);
            blk = prefix[gi] - prefix[s.sub_base];
            const int dY = (int)(short)(prefix[n_subs_total + gi] - prefix[n_subs_total + s.sub_base]);
            const int dB = (int)(short)(prefix[2 * n_subs_total + gi] - prefix[2 * n_subs_total + s.sub_base]);
            const int dR = (int)(short)(prefix[3 * n_subs_total + gi] - prefix[3 * n_subs_total + s.sub_base]);
            const uint32_t comp = (uint32_t)c < ny ? 0u : ((uint32_t)c == ny ? 1u : 2u);
            // rotate so that (p0, t0) belong to the component of block c
            if (comp == 0) { p0 = dY; p1 = dB; p2 = dR; }
            else if (comp == 1) { p0 = dB; p1 = dR; p2 = dY; t0 = sh_tab + 2 * kTabBytes; t1 = sh_tab + 4 * kTabBytes; t2 = sh_tab; }
            else { p0 = dR; p1 = dY; p2 = dB; t0 = sh_tab + 4 * kTabBytes; t1 = sh_tab; t2 = sh_tab + 2 * kTabBytes; }
            owned = (k == 0);
            finished = (p >= end_bit) || (owned && blk >= n_blocks);
            bpos = (uint32_t)(p >> 3);
            hi = ss_load_be32(D, bpos); lo = ss_load_be32(D, bpos + 4);
            bpos += 8;
            const uint32_t sh0 = (uint32_t)p & 7u;
            hi = __funnelshift_l(lo, hi, sh0);
            lo <<= sh0;
            nbits = 64 - (int)sh0;
        }
    }
    const bool last_sub = (li + 1 == s.n_subs) || ((uint64_t)(li + 1) * HJD_SS_SUB_BYTES >= L);

    while (__any_sync(0xffffffffu, !finished)) {
        bool done_block = false;
        if (!finished) {
#pragma unroll
            for (int rep = 0; rep < SS_WRITE_SYMS; rep++) {
                if (!done_block && !finished) {
                    if (nbits < 32) {
                        const uint32_t w = ss_load_be32(D, bpos);
                        bpos += 4;
                        hi |= hjd_shr(w, (uint32_t)nbits);
                        lo |= hjd_shl(w, 32u - (uint32_t)nbits);
                        nbits += 32;
                    }
                    const bool is_ac = k != 0;
                    const SsSym sy = ss_symbol(t0 + (is_ac ? kTabBytes : 0u), hi, lo, is_ac);
                    if (sy.bad) flags |= HJD_ST_BAD_CODE;
                    hi = __funnelshift_l(lo, hi, sy.used);
                    lo <<= sy.used;
                    nbits -= (int)sy.used;
                    p += sy.used;
                    const uint32_t kpos = (uint32_t)k + sy.kadv - 1u;
                    if (owned && sy.size) {
                        if (kpos <= 63u) hjd_sts_u16_sync(my_slot + ((kpos << 1) ^ swz), (uint32_t)sy.val);
                        else flags |= HJD_ST_COEF_RANGE;
                    }
                    k += (int)sy.kadv;
                    if (k >= 64) {
                        if (owned) done_block = true;             // hand-over below
                        else {                                    // the foreign block is over: the next one is mine
                            owned = true;
                            k = 0;
                            if ((uint32_t)++c == bpm) c = 0;
                            if (bpm > 1 && (c == 0 || (uint32_t)c >= ny)) {
                                const int tp = p0; p0 = p1; p1 = p2; p2 = tp;
                                const uint32_t tt = t0; t0 = t1; t1 = t2; t2 = tt;
                            }
                            if (p >= end_bit || blk >= n_blocks) finished = true;
                        }
                    } else if (!owned && p >= end_bit) {
                        finished = true;                          // one block covers this whole sub-sequence
                    }
                }
            }
        }
        // ---- block hand-over ---------------------------------------------------------------
        uint32_t flush_blk = 0;
        if (done_block) {
            p0 = (int)(short)(p0 + (int)(short)hjd_lds_u16_sync(my_slot + swz));
            hjd_sts_u16_sync(my_slot + swz, (uint32_t)p0);
            flush_blk = blk_base + blk;
            blk++;
            k = 0;
            if ((uint32_t)++c == bpm) c = 0;
            if (bpm > 1 && (c == 0 || (uint32_t)c >= ny)) {
                const int tp = p0; p0 = p1; p1 = p2; p2 = tp;
                const uint32_t tt = t0; t0 = t1; t1 = t2; t2 = tt;
            }
            if (p >= end_bit || blk >= n_blocks) finished = true;
        }
        // ---- cooperative flush, four blocks per step (as in kernel 1a) ----------------------
        const uint32_t m = __ballot_sync(0xffffffffu, done_block);
        if (m) {
            if (done_block) hjd_sts_v2_sync(warp_list + (uint32_t)__popc(m & lt_mask) * 8u, flush_blk, (uint32_t)lane);
            __syncwarp();
            const int n_done = __popc(m);
            const uint32_t chunk = (uint32_t)lane & 7u;
            for (int base = 0; base < n_done; base += 4) {
                const int idx = base + (lane >> 3);
                if (idx < n_done) {
                    const uint2 ent = hjd_lds_v2_sync(warp_list + (uint32_t)idx * 8u);
                    const uint32_t src = warp_slots + ent.y * 128u + ((chunk ^ (ent.y & 7u)) << 4);
                    const uint4 w = hjd_lds_v4_sync(src);
                    hjd_sts_zero16_sync(src);
                    ((uint4*)coef)[(size_t)(ent.x * 8u + chunk)] = w;
                }
            }
            __syncwarp();
        }
    }
    // the last sub-sequence that holds data must have completed the image
    if (last_sub && li < s.n_subs && (uint64_t)li * HJD_SS_SUB_BYTES < L && blk < n_blocks) flags |= HJD_ST_OVERRUN;
    if (flags) atomicOr(&status[s.img], flags);
}

cudaError_t hjd_launch_ss_write(const HjdImageDesc* imgs, const HjdTableSet* tsets, const HjdSsImage* ss,
                                const HjdSsWork* work, int n_work, const uint8_t* dst, const uint32_t* dlen,
                                uint32_t n_subs_total, const uint64_t* x, const uint32_t* prefix, int16_t* coef,
                                int32_t* status, cudaStream_t st)
{
    if (n_work <= 0) return cudaSuccess;
    const size_t smem = HJD_SS_THREADS * (128 + 8) + 6 * sizeof(HjdHuffTable);
    hjd_k_ss_write<<<n_work, HJD_SS_THREADS, smem, st>>>(imgs, tsets, ss, work, dst, dlen, n_subs_total, x, prefix,
                                                        coef, status);
    return cudaGetLastError();
}
