// csrc/hjd_pipeline.cpp -- host orchestration on top of the batch C ABI (include/hjd.h):
//
//   * hjd_multi_*: ONE process, several GPUs.  A batch of independent images is cut into one contiguous
//     image range per device, balanced by compressed bytes (the entropy kernel's time follows the scan
//     size); one host thread and one hjd_batch per device, no exchange between them.
//   * hjd_convert_jpg_files*: the reference's ConvertJpgFile (openjpg.cpp:593-684: fopen/fread, decode,
//     WriteBMP24) at batch scale, as a pipeline: parallel file readers fill a pinned arena; per device a
//     few workers each decode one chunk of images at a time with HJD_FLAG_BMP_OUT, so what arrives in
//     their pinned buffers already IS the BMP files (header, bottom-up B G R rows, padding: the colour
//     kernel's epilogue wrote them); writer threads fwrite each file as soon as its chunk has landed, while
//     the other workers' chunks are being copied and decoded.  No CPU pixel pass, no whole-batch buffer.
//
// Everything here uses the public C ABI only; nothing is decoded on the CPU.
#include "../../include/hjd.h"

#include <stdio.h>
#include <string.h>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_perr;

uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

// [lo, hi) image range of `rank`: contiguous, disjoint, covering, balanced by sum(sizes).
// The same rule as hls_jpeg_decoder_b200/sharding.py (tests compare the two).
void shard_range(const int64_t* sizes, int n, int rank, int world, int* lo, int* hi)
{
    if (world <= 1) { *lo = 0; *hi = n; return; }
    long double total = 0;
    for (int i = 0; i < n; i++) total += (long double)(sizes[i] > 0 ? sizes[i] : 0);
    std::vector<int> bounds(1, 0);
    long double acc = 0;
    int k = 1;
    for (int i = 0; i < n; i++) {
        acc += (long double)(sizes[i] > 0 ? sizes[i] : 0);
        while (k < world && acc * world >= total * k) { bounds.push_back(i + 1); k++; }
    }
    while ((int)bounds.size() < world) bounds.push_back(n);
    bounds.push_back(n);
    for (size_t j = 0; j < bounds.size(); j++) { if (bounds[j] > n) bounds[j] = n; if (j && bounds[j] < bounds[j - 1]) bounds[j] = bounds[j - 1]; }
    *lo = bounds[(size_t)rank];
    *hi = bounds[(size_t)rank + 1];
}

} // namespace

struct hjd_multi {
    std::vector<int> devices;
    std::vector<hjd_batch*> batches;
    unsigned flags = 0;
    std::string err;
};

extern "C" int hjd_shard_range(const int64_t* sizes, int n, int rank, int world, int* lo, int* hi)
{
    if (!sizes || n < 0 || world < 1 || rank < 0 || rank >= world || !lo || !hi) return HJD_ERR_ARG;
    shard_range(sizes, n, rank, world, lo, hi);
    return HJD_OK;
}

extern "C" hjd_multi* hjd_multi_create(const int* devices, int n_devices, unsigned flags)
{
    const int have = hjd_device_count();
    if (n_devices <= 0 || have <= 0) return nullptr;
    hjd_multi* m = new hjd_multi();
    m->flags = flags;
    for (int k = 0; k < n_devices; k++) {
        const int dev = devices ? devices[k] : k;
        hjd_batch* b = (dev >= 0 && dev < have) ? hjd_batch_create(dev, flags) : nullptr;
        if (!b) { hjd_multi_destroy(m); return nullptr; }
        m->devices.push_back(dev);
        m->batches.push_back(b);
    }
    return m;
}

extern "C" void hjd_multi_destroy(hjd_multi* m)
{
    if (!m) return;
    for (hjd_batch* b : m->batches) hjd_batch_destroy(b);
    delete m;
}

extern "C" int hjd_multi_num_devices(const hjd_multi* m) { return m ? (int)m->batches.size() : 0; }
extern "C" hjd_batch* hjd_multi_batch(hjd_multi* m, int k) { return (m && k >= 0 && k < (int)m->batches.size()) ? m->batches[(size_t)k] : nullptr; }
extern "C" const char* hjd_multi_last_error(const hjd_multi* m) { return m ? m->err.c_str() : "null handle"; }

extern "C" int hjd_multi_decode_host(hjd_multi* m, const uint8_t* arena, const int64_t* offsets, const int64_t* sizes, int n,
                                     uint8_t* rgb_out, uint64_t rgb_capacity, uint64_t* rgb_offsets_out, int32_t* status_out)
{
    if (!m || !arena || !offsets || !sizes || !rgb_out || n < 0) return HJD_ERR_ARG;
    const int world = (int)m->batches.size();
    std::vector<int> lo((size_t)world), hi((size_t)world);
    for (int k = 0; k < world; k++) shard_range(sizes, n, k, world, &lo[(size_t)k], &hi[(size_t)k]);
    // phase 1: every shard sizes its part of the output (header parse only), in parallel
    std::vector<uint64_t> need((size_t)world, 0);
    {
        std::vector<std::thread> pool;
        for (int k = 0; k < world; k++)
            pool.emplace_back([&, k]() {
                need[(size_t)k] = hjd_out_slab_bytes(arena, offsets + lo[(size_t)k], sizes + lo[(size_t)k], hi[(size_t)k] - lo[(size_t)k], m->flags);
            });
        for (auto& t : pool) t.join();
    }
    std::vector<uint64_t> base((size_t)world + 1, 0);
    for (int k = 0; k < world; k++) base[(size_t)k + 1] = base[(size_t)k] + need[(size_t)k];
    if (base[(size_t)world] > rgb_capacity) { m->err = "rgb_out too small"; return HJD_ERR_ARG; }
    // phase 2: one host thread per device: upload, decode, download of its own image range
    std::vector<int> rc((size_t)world, HJD_OK);
    std::vector<std::string> errs((size_t)world);
    {
        std::vector<std::thread> pool;
        for (int k = 0; k < world; k++)
            pool.emplace_back([&, k]() {
                const int a = lo[(size_t)k], cnt = hi[(size_t)k] - a;
                if (cnt <= 0) return;
                rc[(size_t)k] = hjd_batch_decode_host(m->batches[(size_t)k], arena, offsets + a, sizes + a, cnt,
                                                      rgb_out + base[(size_t)k], need[(size_t)k],
                                                      rgb_offsets_out ? rgb_offsets_out + a : nullptr,
                                                      status_out ? status_out + a : nullptr, 0);
                if (rc[(size_t)k] != HJD_OK) errs[(size_t)k] = hjd_last_error();
                else if (rgb_offsets_out)
                    for (int i = 0; i < cnt; i++) rgb_offsets_out[a + i] += base[(size_t)k];
            });
        for (auto& t : pool) t.join();
    }
    for (int k = 0; k < world; k++)
        if (rc[(size_t)k] != HJD_OK) { m->err = "device " + std::to_string(m->devices[(size_t)k]) + ": " + errs[(size_t)k]; return rc[(size_t)k]; }
    return HJD_OK;
}

extern "C" uint64_t hjd_multi_out_slab_bytes(const hjd_multi* m, const uint8_t* arena, const int64_t* offsets, const int64_t* sizes, int n)
{
    if (!m || !arena || !offsets || !sizes || n < 0) return 0;
    const int world = (int)m->batches.size();
    uint64_t total = 0;
    for (int k = 0; k < world; k++) {
        int lo, hi;
        shard_range(sizes, n, k, world, &lo, &hi);
        total += hjd_out_slab_bytes(arena, offsets + lo, sizes + lo, hi - lo, m->flags);
    }
    return total;
}

// ------------------------------------------------------------------------------------------
// file -> BMP pipeline
// ------------------------------------------------------------------------------------------
namespace {

struct WriteTask { const uint8_t* src; size_t bytes; const char* path; int* ok; std::atomic<int>* pending; std::atomic<int>* converted; };

struct WriterPool {
    std::mutex mu;
    std::condition_variable cv;
    std::deque<WriteTask> q;
    bool closing = false;
    std::vector<std::thread> threads;

    void start(int n)
    {
        for (int t = 0; t < n; t++) threads.emplace_back([this]() { run(); });
    }
    void push(const WriteTask& w)
    {
        { std::lock_guard<std::mutex> l(mu); q.push_back(w); }
        cv.notify_one();
    }
    void run()
    {
        for (;;) {
            WriteTask w;
            {
                std::unique_lock<std::mutex> l(mu);
                cv.wait(l, [this]() { return closing || !q.empty(); });
                if (q.empty()) return;
                w = q.front();
                q.pop_front();
            }
            int good = 0;
            FILE* fp = fopen(w.path, "wb");
            if (fp) {
                good = fwrite(w.src, 1, w.bytes, fp) == w.bytes;
                if (fclose(fp) != 0) good = 0;
            }
            if (good) { if (w.ok) *w.ok = 1; w.converted->fetch_add(1); }
            w.pending->fetch_sub(1);
        }
    }
    void stop()
    {
        { std::lock_guard<std::mutex> l(mu); closing = true; }
        cv.notify_all();
        for (auto& t : threads) t.join();
    }
};

} // namespace

extern "C" int hjd_convert_jpg_files_multi(const char* const* jpg_in, const char* const* bmp_out, int n,
                                           const int* devices, int n_devices, int threads, int chunk_images, int* ok)
{
    if (!jpg_in || !bmp_out || n < 0 || n_devices <= 0) return 0;
    if (ok) for (int i = 0; i < n; i++) ok[i] = 0;
    if (n == 0) return 0;
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = 1;
    if (chunk_images <= 0) chunk_images = 32;
    const int n_chunks = (n + chunk_images - 1) / chunk_images;

    // 1. sizes (the batched counterpart of FileSize, openjpg.cpp:579-586), in parallel
    std::vector<int64_t> sizes((size_t)n, 0), offsets((size_t)n, 0);
    {
        std::atomic<int> next(0);
        auto stat_one = [&]() {
            for (int i = next++; i < n; i = next++) {
                FILE* fp = jpg_in[i] ? fopen(jpg_in[i], "rb") : nullptr;
                if (fp) { fseek(fp, 0, SEEK_END); long len = ftell(fp); fclose(fp); if (len > 0) sizes[(size_t)i] = len; }
            }
        };
        std::vector<std::thread> pool;
        for (int t = 0; t < (threads < n ? threads : n); t++) pool.emplace_back(stat_one);
        for (auto& t : pool) t.join();
    }
    int64_t total = 0;
    for (int i = 0; i < n; i++) { offsets[(size_t)i] = total; total += (int64_t)align_up((uint64_t)sizes[(size_t)i], 16); }
    uint8_t* arena = (uint8_t*)hjd_host_alloc((size_t)total + 16);
    if (!arena) return 0;

    // 2. readers: chunk by chunk, in chunk order, so that the first decode can start before the last file is read
    std::vector<std::atomic<int>> chunk_left((size_t)n_chunks);
    for (int c = 0; c < n_chunks; c++) chunk_left[(size_t)c] = (c + 1 < n_chunks ? chunk_images : n - c * chunk_images);
    std::mutex ready_mu;
    std::condition_variable ready_cv;
    std::atomic<int> next_file(0);
    auto reader = [&]() {
        for (int i = next_file++; i < n; i = next_file++) {
            if (sizes[(size_t)i]) {                                                       // openjpg.cpp:603-619
                FILE* fp = fopen(jpg_in[i], "rb");
                size_t got = fp ? fread(arena + offsets[(size_t)i], 1, (size_t)sizes[(size_t)i], fp) : 0;
                if (fp) fclose(fp);
                if (got != (size_t)sizes[(size_t)i]) sizes[(size_t)i] = 0;
            }
            if (chunk_left[(size_t)(i / chunk_images)].fetch_sub(1) == 1) { std::lock_guard<std::mutex> l(ready_mu); ready_cv.notify_all(); }
        }
    };
    std::vector<std::thread> readers;
    for (int t = 0; t < (threads < n ? threads : n); t++) readers.emplace_back(reader);

    // 3. writers + per-device decode workers
    WriterPool writers;
    writers.start(threads);
    std::atomic<int> converted(0), next_chunk(0);
    const int workers_per_device = n_chunks >= 3 * n_devices ? 3 : (n_chunks >= 2 * n_devices ? 2 : 1);
    auto worker = [&](int device) {
        hjd_batch* b = hjd_batch_create(device, HJD_FLAG_BMP_OUT);
        if (!b) return;
        uint8_t* out = nullptr;
        uint64_t cap = 0;
        std::vector<uint64_t> offs((size_t)chunk_images);
        std::vector<int32_t> status((size_t)chunk_images);
        std::atomic<int> pending(0);
        for (int c = next_chunk++; c < n_chunks; c = next_chunk++) {
            const int first = c * chunk_images, cnt = (c + 1 < n_chunks ? chunk_images : n - first);
            {
                std::unique_lock<std::mutex> l(ready_mu);
                ready_cv.wait(l, [&]() { return chunk_left[(size_t)c].load() <= 0; });
            }
            const uint64_t need = hjd_out_slab_bytes(arena, offsets.data() + first, sizes.data() + first, cnt, HJD_FLAG_BMP_OUT);
            while (pending.load() > 0) std::this_thread::yield();                        // the previous chunk's files are on disk
            if (need > cap) {
                if (out) hjd_host_free(out);
                cap = need + need / 8 + 4096;
                out = (uint8_t*)hjd_host_alloc((size_t)cap);
                if (!out) { cap = 0; continue; }
            }
            if (hjd_batch_decode_host(b, arena, offsets.data() + first, sizes.data() + first, cnt, out, cap, offs.data(),
                                      status.data(), 0) != HJD_OK)
                continue;
            for (int i = 0; i < cnt; i++) {
                const uint64_t bytes = hjd_batch_bmp_bytes(b, i);
                if (status[(size_t)i] < 0 || !bytes || !bmp_out[first + i]) continue;   // rejected by the parser
                pending.fetch_add(1);
                writers.push(WriteTask{out + offs[(size_t)i] + 10, (size_t)bytes, bmp_out[first + i], ok ? ok + first + i : nullptr,
                                       &pending, &converted});
            }
        }
        while (pending.load() > 0) std::this_thread::yield();
        if (out) hjd_host_free(out);
        hjd_batch_destroy(b);
    };
    std::vector<std::thread> workers;
    for (int k = 0; k < n_devices; k++)
        for (int w = 0; w < workers_per_device; w++) workers.emplace_back(worker, devices ? devices[k] : k);
    for (auto& t : readers) t.join();
    { std::lock_guard<std::mutex> l(ready_mu); ready_cv.notify_all(); }
    for (auto& t : workers) t.join();
    writers.stop();
    hjd_host_free(arena);
    return converted.load();
}

extern "C" int hjd_convert_jpg_files(const char* const* jpg_in, const char* const* bmp_out, int n, int device,
                                     int threads, int* ok)
{
    return hjd_convert_jpg_files_multi(jpg_in, bmp_out, n, &device, 1, threads, 0, ok);
}
