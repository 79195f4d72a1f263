// csrc/selfsync.cuh -- kernel 1b: speculative self-synchronising Huffman decode of restart-free scans.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "hjd_types.h"

// Device-wide exclusive prefix sum of n uint32 (in place in `data`); tmp holds >= n/2048 + 2 words.
cudaError_t hjd_scan_u32(uint32_t* data, uint32_t n, uint32_t* tmp, cudaStream_t st);

// De-stuffing pre-pass (FillNBits' FF00 rule, loadjpg.cpp:475-478, applied once, in parallel):
// The n_ss images given own the 16-byte chunks [first_chunk, first_chunk + n_chunks) of the batch-wide numbering
// (HjdSsImage::chunk_base); counts[] is scratch over that range + one sentinel; dst receives the compacted
// streams, dlen[k] the length of ss[k]'s.
cudaError_t hjd_launch_destuff(const uint8_t* arena, const HjdImageDesc* imgs, const HjdSsImage* ss, int n_ss,
                               uint32_t first_chunk, uint32_t n_chunks, uint32_t* counts, uint32_t* scan_tmp,
                               uint8_t* dst, uint32_t* dlen, cudaStream_t st);

// Speculative decode of every sub-sequence (one CTA per `work` entry, HJD_SS_THREADS sub-sequences).
// e: exit states, x: the entry state each exit was computed from, cnt: [4][n_subs_total] blocks
// started / DC-difference sums per component of each sub-sequence.
cudaError_t hjd_launch_ss_spec(const HjdImageDesc* imgs, const HjdTableSet* tsets, const HjdSsImage* ss,
                               const HjdSsWork* work, const HjdSsSeg* segs, int n_work, const uint8_t* dst,
                               const uint32_t* dlen,
                               uint32_t n_subs_total, uint64_t* e, uint64_t* x, uint32_t* cnt, cudaStream_t st);

// The synchronisation rounds, in place, all of them in one persistent cooperative kernel (grid-wide barrier
// between rounds, termination decided on the device).  `work` here has at most HJD_SS_FIX_WARPS segments per
// entry, one range (<= HJD_SS_FIX_MAXR sub-sequences) per warp.  ctl: three zeroed words (barrier count,
// last round with work, rounds executed | bit 31 if max_rounds was hit).  max_resident_ctas: from
// hjd_selfsync_init_device.
cudaError_t hjd_launch_ss_sync(const HjdImageDesc* imgs, const HjdTableSet* tsets, const HjdSsImage* ss,
                               const HjdSsWork* work, const HjdSsSeg* segs, int n_work, const uint8_t* dst,
                               const uint32_t* dlen,
                               uint32_t n_subs_total, uint64_t* e, uint64_t* x, uint32_t* cnt,
                               uint32_t* ctl, uint32_t max_rounds, int max_resident_ctas, cudaStream_t st);

// Per-device function attributes (call after cudaSetDevice, once per batch handle); *max_sync_ctas = grid
// limit of the cooperative synchronisation kernel on this device.
cudaError_t hjd_selfsync_init_device(int* max_sync_ctas);

// Final pass: every thread decodes, from its (now correct) entry state, the blocks that start in its
// sub-sequence and writes them as whole 128-byte lines, DC un-differenced.  prefix = exclusive scan of cnt.
cudaError_t hjd_launch_ss_write(const HjdImageDesc* imgs, const HjdTableSet* tsets, const HjdSsImage* ss,
                                const HjdSsWork* work, const HjdSsSeg* segs, int n_work, const uint8_t* dst,
                                const uint32_t* dlen,
                                uint32_t n_subs_total, const uint64_t* x, const uint32_t* prefix, int16_t* coef,
                                int32_t* status, cudaStream_t st);

// Zero the blocks of truncated scans that no sub-sequence started (prefix as for hjd_launch_ss_write).
cudaError_t hjd_launch_ss_fill_tail(const HjdImageDesc* imgs, const HjdSsImage* ss, int n_ss, const uint32_t* prefix,
                                    int16_t* coef, cudaStream_t st);
