/* include/hjd.h -- C ABI of the B200-native baseline-JPEG decode path.
 *
 * Drop-in boundary for the hot path of harutel/hls-jpeg-decoder.  The
 * reference has plain C++ linkage and no FFI; these are the entry points a binding (cgo,
 * JNI, ctypes, ...) or the reference's own main.cpp would bind instead of its CPU decode:
 *
 *   reference interface (file:line)                           -> replacement
 *   ---------------------------------------------------------------------------------------
 *   int  ConvertJpgFile(char*, char*)        openjpg.h:23, openjpg.cpp:593  -> hjd_convert_jpg_file
 *   int  DecodeJpgFileData(buf,size,&rgb,&w,&h)  loadjpg.h:186 (declared only) -> hjd_decode_jpg_file_data
 *   void JpegGetImageSize(..., &w, &h)       loadjpg.h:183 (declared only)  -> hjd_get_image_size
 *   void WriteBMP24(name, w, h, rgb)         openjpg.cpp:504                -> hjd_write_bmp24
 *   int  JpegDecodeHW(stJpegData*, h, w, hF, vF) loadjpg.h:180, loadjpg.cpp:1134
 *        (the HLS top: parsed tables + entropy segment in, RGB out)         -> hjd_batch_* (N images per call)
 *
 * csrc/ref_shim.cpp additionally defines the reference's exact C++ names on top of this ABI
 * so that the reference's src/main.cpp links unchanged.
 *
 * Conventions: plain pointers and sizes only; no global mutable state besides a lazily
 * created per-thread default batch used by the single-image calls; one hjd_batch per GPU,
 * used from one host thread at a time.  All hot-path work runs in hand-written sm_100a CUDA
 * kernels; there is no CPU fallback: without a usable GPU every decode call fails with
 * HJD_ERR_CUDA and hjd_last_error() says why.
 *
 * Output contract (loadjpg.cpp:921-925): RGB24, top-down, tightly packed, stride 3*width, R first.
 * BMP contract (openjpg.cpp:504-570): 54-byte header, bottom-up rows, B,G,R, rows padded to 4 bytes.
 */
#ifndef HJD_H
#define HJD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HJD_VERSION 100

/* Return codes of the batch API (single-image reference-style calls return 1/0, see below). */
#define HJD_OK                 0
#define HJD_ERR_ARG           -1
#define HJD_ERR_CUDA          -2   /* no device / CUDA runtime failure; see hjd_last_error() */
#define HJD_ERR_NOMEM         -3
#define HJD_ERR_IO            -4
#define HJD_ERR_STATE         -5   /* call order (e.g. decode before upload) */

/* Per-image status (hjd_batch_get_status).  0 = decoded; parse errors are negative and the
 * image is skipped (its RGB slab is left zero-filled); decode warnings are positive bit flags
 * and the image is still produced (the reference itself only printf()s on such input). */
#define HJD_IMG_OK                 0
#define HJD_IMG_ERR_NOT_JPEG      -1   /* openjpg.cpp:481-486 */
#define HJD_IMG_ERR_TRUNCATED     -2
#define HJD_IMG_ERR_UNSUPPORTED   -3   /* progressive, 12-bit, 16-bit DQT, CMYK, chroma sampling != 1x1 ... */
#define HJD_IMG_ERR_BAD_TABLE     -4   /* DHT over-subscribed / missing table */
#define HJD_IMG_WARN_BAD_CODE      1   /* undecodable Huffman code; rest of that restart interval zero-filled */
#define HJD_IMG_WARN_COEF_RANGE    2   /* run past coefficient 63 (loadjpg.cpp:780-783) */
#define HJD_IMG_WARN_OVERRUN       4   /* entropy data ended before the interval did */
#define HJD_IMG_WARN_RESTART       8   /* RSTn count differs from ceil(MCUs/Ri)-1 */

/* hjd_batch_create flags */
#define HJD_FLAG_KEEP_PLANES   1u   /* unfused kernels 2 and 3 with the Y/Cb/Cr planes in HBM (parity tap, hjd_batch_download_planes) */
#define HJD_FLAG_HOST_SCAN     2u   /* find RSTn markers on the host instead of the GPU pre-pass */
#define HJD_FLAG_FUSED_MCU    16u   /* (default behaviour) kernels 2+3 fused per MCU: one thread decodes a whole MCU to RGB */
#define HJD_FLAG_BMP_OUT      32u   /* the output slab holds, per image, the BMP FILE WriteBMP24 would write (openjpg.cpp:504-570:
                                       54-byte header, bottom-up B G R rows padded to 4 bytes) instead of top-down RGB24: written by
                                       the colour kernel's epilogue, so a file writer only has to fwrite it.  The file of image i
                                       is hjd_batch_bmp_bytes(i) bytes from rgb_offset + 10 of the slab (pixel array 64-byte aligned). */
#define HJD_FLAG_TENSOR_CORE_IDCT 128u /* always the fused kernel with the IDCT's fast tier as tcgen05.mma on the tensor cores (csrc/mcu_tc.cuh) */
#define HJD_FLAG_CUDA_CORE_IDCT   64u  /* always the fused kernel with the fast tier as FP32 FMA chains on the CUDA cores (csrc/kernels.cu);
                                          default: chosen per chunk -- tensor cores for large colour images, CUDA cores for batches of small ones (same bytes out) */
#define HJD_FLAG_NO_SELFSYNC   8u   /* restart-free scans: one thread per scan (kernel 1a) instead of kernel 1b */

typedef struct hjd_batch hjd_batch;

typedef struct hjd_image_info {
    uint32_t width, height;
    uint8_t  ncomp;            /* 1 or 3 */
    uint8_t  hf, vf;           /* luma sampling factors (1 or 2); chroma is 1x1 */
    uint8_t  blocks_per_mcu;
    uint32_t mcus_x, mcus_y;
    uint32_t restart_interval; /* MCUs, 0 = none */
    uint32_t n_intervals;
    uint32_t scan_bytes;       /* entropy-coded segment length */
    uint64_t block_base;       /* first block of this image in the coefficient buffer */
    uint64_t n_blocks;
    uint64_t rgb_offset;       /* byte offset of this image in the RGB slab */
    uint64_t y_offset, cb_offset, cr_offset;   /* byte offsets in the plane slab */
    uint32_t y_pitch, c_pitch;
    int32_t  status;
} hjd_image_info;

/* Stage timings of the last hjd_batch_decode (CUDA events on the batch stream), milliseconds. */
typedef struct hjd_timings {
    float scan_ms;      /* kernel 0: RSTn marker scan -> interval table */
    float entropy_ms;   /* kernel 1: Huffman decode -> int16 coefficients (+ self-sync passes) */
    float idct_ms;      /* kernel 2 (or fused 2+3) */
    float color_ms;     /* kernel 3 (0 when fused) */
    float total_ms;
    int   launches;     /* kernels launched by the last decode */
} hjd_timings;

/* ---- library ---------------------------------------------------------------------------- */
int         hjd_version(void);
const char* hjd_last_error(void);          /* thread-local, never NULL */
int         hjd_device_count(void);        /* 0 when no CUDA device is usable */

/* ---- reference-shaped single-image entry points (return 1 = success, 0 = failure) -------- */
/* openjpg.cpp:593 ConvertJpgFile: load .jpg, decode on the GPU, write 24-bit .bmp. */
int  hjd_convert_jpg_file(const char* jpg_in, const char* bmp_out);
/* loadjpg.h:186 DecodeJpgFileData: *rgb is allocated by the library (release with hjd_free). */
int  hjd_decode_jpg_file_data(const uint8_t* buf, int size, uint8_t** rgb, unsigned* width, unsigned* height);
void hjd_free(void* p);
/* Same, with the result allocated by the caller's allocator (the C++ shim passes operator new[], so that the
 * reference's "delete[] rgbpix" contract holds without a second copy). */
int  hjd_decode_jpg_file_data_alloc(const uint8_t* buf, int size, void* (*alloc)(size_t), uint8_t** rgb,
                                    unsigned* width, unsigned* height);
/* Device used by the reference-shaped single-image calls (a lazily created per-thread batch); default 0. */
int  hjd_set_default_device(int device);
/* loadjpg.h:183 JpegGetImageSize, from the file bytes (header parse only, no GPU). */
int  hjd_get_image_size(const uint8_t* buf, int size, unsigned* width, unsigned* height);
/* Header parse only (no GPU): the per-image status a batch would report for this file (HJD_IMG_OK or a
 * negative HJD_IMG_ERR_*) and, when it is decodable, its geometry (offsets in *out stay 0).
 * Not decodable, by design: progressive / lossless / arithmetic (SOF2..SOF15), 12-bit samples (SOF1 with
 * P != 8), 16-bit DQT, 4 components, chroma sampling other than 1x1, luma factors other than 1 or 2, and
 * NON-INTERLEAVED baseline files (one scan per component): the reference's decode loop is one interleaved
 * scan of all three components (loadjpg.cpp:945-997), so such files get HJD_IMG_ERR_UNSUPPORTED. */
int  hjd_probe_jpeg(const uint8_t* buf, int64_t size, hjd_image_info* out);
/* openjpg.cpp:504 WriteBMP24. */
int  hjd_write_bmp24(const char* path, unsigned width, unsigned height, const uint8_t* rgb);
/* Same bytes as hjd_write_bmp24 into memory; returns the BMP size (call with out = NULL to size it). */
size_t hjd_encode_bmp24(unsigned width, unsigned height, const uint8_t* rgb, uint8_t* out);

/* ConvertJpgFile at batch scale, as a pipeline (openjpg.cpp:593-684 per file): `threads` readers fill a pinned
 * arena; per device up to three workers decode one chunk of `chunk_images` images at a time with
 * HJD_FLAG_BMP_OUT (the GPU writes the BMP file layout); `threads` writers fwrite every file as soon as its
 * chunk has landed in pinned memory, while later chunks are being copied and decoded (0 = all cores / 32
 * images).  ok[i] = 1 / 0 per file (may be NULL).  Returns the number of files converted. */
int  hjd_convert_jpg_files(const char* const* jpg_in, const char* const* bmp_out, int n, int device,
                           int threads, int* ok);
int  hjd_convert_jpg_files_multi(const char* const* jpg_in, const char* const* bmp_out, int n,
                                 const int* devices, int n_devices, int threads, int chunk_images, int* ok);

/* ---- batch API: the JpegDecodeHW replacement, N independent images per call -------------- */
hjd_batch* hjd_batch_create(int device, unsigned flags);
void       hjd_batch_destroy(hjd_batch* b);
/* Use an existing CUDA stream (cudaStream_t passed as void*) instead of the batch's own. */
int  hjd_batch_set_stream(hjd_batch* b, void* cuda_stream);

/* Parse headers, build Huffman/quant tables and descriptors, copy the files to HBM.
 * bufs[i]/sizes[i]: whole .jpg files in host memory (pinned memory makes the copy asynchronous). */
int  hjd_batch_upload(hjd_batch* b, const uint8_t* const* bufs, const int64_t* sizes, int n);
/* Same, for files packed in one host arena: file i = arena[offsets[i] .. offsets[i]+sizes[i]).
 * One host->device copy of [offsets[0], offsets[n-1]+sizes[n-1]). */
int  hjd_batch_upload_arena(hjd_batch* b, const uint8_t* arena, const int64_t* offsets, const int64_t* sizes, int n);

/* Launch the decode of everything uploaded (asynchronous on the batch stream):
 * marker scan -> entropy decode -> dequant/IDCT -> upsample/colour.  Results stay in HBM. */
int  hjd_batch_decode(hjd_batch* b);
int  hjd_batch_sync(hjd_batch* b);
/* hjd_batch_decode_host splits large batches into up to 8 chunks on 3 streams so that the H2D copy,
 * the kernels and the D2H copy of different chunks overlap (default, on = 1).  HBM-resident decodes
 * (hjd_batch_upload* + hjd_batch_decode) run on one stream, in order: measured on B200, overlapping
 * their kernels gains nothing (all are issue-bound) and per-stage timings need the serial order.
 * on = 0: never chunk; on > 1: chunk both paths with this target number of 8x8 blocks per chunk
 * (testing / tuning).  Takes effect at the next upload. */
int  hjd_batch_set_overlap(hjd_batch* b, int on);

/* Synchronisation rounds the self-synchronising kernel (restart-free scans) needed in the last decode. */
int  hjd_batch_selfsync_rounds(hjd_batch* b);                        /* syncs */
/* Sub-sequences per warp in those rounds: 0 = pick by batch size (default), else a multiple of 32 up
 * to 256 (testing / tuning; results do not depend on it).  Takes effect at the next upload. */
int  hjd_batch_set_selfsync_range(hjd_batch* b, int range);
/* Which fused kernel(s) the chunks of the last decode went through: bit 0 = tensor cores (csrc/mcu_tc.cuh), bit 1 = CUDA
 * cores; 0 = none (nothing decoded yet, or HJD_FLAG_KEEP_PLANES). */
int  hjd_batch_idct_variant(const hjd_batch* b);
int  hjd_batch_num_images(const hjd_batch* b);
int  hjd_batch_get_info(const hjd_batch* b, int i, hjd_image_info* out);
int  hjd_batch_get_status(hjd_batch* b, int32_t* status /* n */);   /* syncs */
int  hjd_batch_get_timings(hjd_batch* b, hjd_timings* out);          /* syncs */
/* CUDA-event stopwatch on the batch stream: record event `slot` (0..HJD_MARK_SLOTS-1); milliseconds
 * between two recorded slots (waits for slot_b).  This is how bench.py times K steps, and each of
 * them, on the launching stream. */
#define HJD_MARK_SLOTS 130
int   hjd_batch_mark(hjd_batch* b, int slot);
float hjd_batch_elapsed_ms(hjd_batch* b, int slot_a, int slot_b);
uint64_t hjd_batch_rgb_bytes(const hjd_batch* b);     /* size of the RGB slab */
uint64_t hjd_batch_coef_bytes(const hjd_batch* b);
uint64_t hjd_batch_plane_bytes(const hjd_batch* b);
uint64_t hjd_batch_scan_bytes(const hjd_batch* b);    /* sum of entropy-coded bytes */
uint64_t hjd_batch_pixels(const hjd_batch* b);        /* sum of width*height */

/* Device pointers of the result slabs (valid until the next upload / destroy). */
void* hjd_batch_device_rgb(hjd_batch* b);
void* hjd_batch_device_coef(hjd_batch* b);
void* hjd_batch_device_planes(hjd_batch* b);

/* Device -> host copies (synchronous with respect to the batch stream). */
int  hjd_batch_download_rgb(hjd_batch* b, uint8_t* dst /* hjd_batch_rgb_bytes */);
int  hjd_batch_download_image(hjd_batch* b, int i, uint8_t* dst /* w*h*3 */);
int  hjd_batch_download_coef(hjd_batch* b, int16_t* dst /* hjd_batch_coef_bytes */);
uint64_t hjd_batch_bmp_bytes(const hjd_batch* b, int i);            /* HJD_FLAG_BMP_OUT: size of image i's BMP file (0: none) */
int  hjd_batch_download_bmp(hjd_batch* b, int i, uint8_t* dst /* hjd_batch_bmp_bytes(i) */);
int  hjd_batch_download_image_coef(hjd_batch* b, int i, int16_t* dst /* n_blocks*64 */);   /* one image of the slab */
int  hjd_batch_download_planes(hjd_batch* b, uint8_t* dst /* hjd_batch_plane_bytes */);

/* End to end with host buffers: upload, decode and download in overlapped chunks
 * (copy engines and SMs busy at the same time).  rgb_out receives image i at
 * rgb_offsets_out[i] (tightly packed, 256-byte aligned starts); capacity in bytes. */
int  hjd_batch_decode_host(hjd_batch* b, const uint8_t* arena, const int64_t* offsets, const int64_t* sizes,
                           int n, uint8_t* rgb_out, uint64_t rgb_capacity, uint64_t* rgb_offsets_out,
                           int32_t* status_out, int chunk_images);
/* Bytes hjd_batch_decode_host needs in rgb_out for these files (header parse only); hjd_out_slab_bytes: the
 * same for a batch created with `flags` (HJD_FLAG_BMP_OUT changes the layout). */
uint64_t hjd_rgb_slab_bytes(const uint8_t* arena, const int64_t* offsets, const int64_t* sizes, int n);
uint64_t hjd_out_slab_bytes(const uint8_t* arena, const int64_t* offsets, const int64_t* sizes, int n, unsigned flags);

/* ---- several GPUs in one process -------------------------------------------
 * Images are independent and there is no exchange step: a batch is cut into one contiguous image range per
 * device, balanced by compressed bytes (hjd_shard_range), and every device decodes its range with its own
 * hjd_batch, driven by its own host thread.  rgb_out receives the shards back to back: image i is at
 * rgb_offsets_out[i], exactly the bytes a single-device decode produces for it. */
typedef struct hjd_multi hjd_multi;
int        hjd_shard_range(const int64_t* sizes, int n, int rank, int world, int* lo, int* hi);
hjd_multi* hjd_multi_create(const int* devices /* NULL: 0..n-1 */, int n_devices, unsigned flags);
void       hjd_multi_destroy(hjd_multi* m);
int        hjd_multi_num_devices(const hjd_multi* m);
hjd_batch* hjd_multi_batch(hjd_multi* m, int k);      /* the k-th device's batch handle (resident use, timings) */
const char* hjd_multi_last_error(const hjd_multi* m);
uint64_t   hjd_multi_out_slab_bytes(const hjd_multi* m, const uint8_t* arena, const int64_t* offsets, const int64_t* sizes, int n);
int        hjd_multi_decode_host(hjd_multi* m, const uint8_t* arena, const int64_t* offsets, const int64_t* sizes, int n,
                                 uint8_t* rgb_out, uint64_t rgb_capacity, uint64_t* rgb_offsets_out, int32_t* status_out);

/* Pinned host memory helpers (cudaHostAlloc, portable: pinned for every device of the process / cudaFreeHost) for callers without a CUDA binding. */
void* hjd_host_alloc(size_t bytes);
void  hjd_host_free(void* p);
/* Same, with the pages taken from the NUMA node next to `device` where the platform exposes one
 * (hjd_device_numa_node: -1 otherwise): the D2H copy of the decoded batch is the end-to-end bottleneck. */
void* hjd_host_alloc_near(int device, size_t bytes);
int   hjd_device_numa_node(int device);
/* Raw host-link probe, no kernels: reps x (h2d_bytes up || d2h_bytes down, pinned buffers, two streams);
 * milliseconds per repetition and direction.  The ceiling hjd_batch_decode_host works against. */
int   hjd_link_probe(int device, const void* host_in, size_t h2d_bytes, void* host_out, size_t d2h_bytes,
                     int reps, float* ms_h2d, float* ms_d2h);

/* Testing aid: build the kernels' two-level lookup table from BITS / HUFFVAL (as in a DHT segment) and look
 * up the next 16 bits of a stream: returns len | size << 5 | zig-zag advance << 9 (0xFE01: no such code;
 * 0xFFFFFFFF: the table is over-subscribed).  tests/ compare it with the canonical code walk. */
uint32_t hjd_huff_lookup_probe(const uint8_t bits[16], const uint8_t* vals, int nvals, int is_ac, uint32_t peek16);

/* The float constants the kernels use (computed on the host with the libm expressions of
 * loadjpg.cpp:96-102,120): cos_tab[p*8+k] = cosf(((2p+1)*k*3.14f)/16), cc[u*8+v] = C(u)*C(v). */
void hjd_get_idct_tables(float cos_tab[64], float cc[64]);
/* The IDCT matrix of the tensor-core kernel as the 16 KB FP16 tile image it multiplies by (csrc/mcu_tc.cuh): with
 * M[k][8y+x] = 0.25 * cc * cos[x][u] * cos[y][v] for zig-zag position k = (u, v) (the real-number product of the constants
 * above), row n of 128 holds M_hi * 2^-13 (n = 8y+x) or M_lo (n = 64 + 8y+x), integers with M = (M_hi * 2^11 + M_lo) * 2^-24
 * up to 2^-25; 64 FP16 per row in zig-zag order, 16-byte chunk j of row n stored at chunk j ^ (n & 7), rows in 1024-byte
 * groups of eight (the K-major, 128-byte-swizzled operand layout of tcgen05.mma).  Host computation, no GPU needed. */
void hjd_get_idct_matrix(uint16_t img[8192]);

#ifdef __cplusplus
}
#endif
#endif /* HJD_H */
