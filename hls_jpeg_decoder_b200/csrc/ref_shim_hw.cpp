// csrc/ref_shim_hw.cpp -- JpegDecodeHW (the reference's HLS top, loadjpg.h:180 / loadjpg.cpp:1134)
// on top of the C ABI, for programs that keep the reference's own parser.
//
// Build it TOGETHER WITH the reference's unchanged src/main.cpp and src/openjpg.cpp (and instead of
// src/loadjpg.cpp), with -I<reference>/src: it needs the reference's stJpegData definition, which
// is why it is not part of libhjd.so.  oracle/build_ref.sh does exactly that to produce the drop-in
// demonstration binary oracle/_ref/ref_main_on_gpu.
//
// The caller hands over what openjpg.cpp parsed: quantisation tables as floats in zig-zag order
// (stComponent::m_qTable), Huffman tables as canonical (code, length, value) lists
// (stHuffmanTable::m_blocks), table selectors and the entropy-coded segment (m_stream).  The adapter
// re-serialises those into a minimal baseline JFIF stream in memory and lets the GPU path decode it;
// RGB lands in jdata->m_rgb exactly as the reference leaves it (loadjpg.cpp:921-925).
// Restart intervals cannot be expressed through this struct: the reference stores Lr, not Ri, in
// m_restart_interval (openjpg.cpp:441-446), so like the reference itself this entry point is for
// restart-free files; use the file/memory entry points for everything else.
//
// JpegGetImageSize(stJpegData*, unsigned*, unsigned*) (loadjpg.h:183) is declared by the reference and never
// defined; its stJpegData carries no dimensions (they live in stImageInfo, openjpg.h:11-17), so the only
// meaning the signature can have is implemented here: the width and height this jdata was last decoded
// with by JpegDecodeHW (0 x 0 if it never was).
#include "loadjpg.h"
#include "../../include/hjd.h"
#include <string.h>
#include <mutex>
#include <unordered_map>
#include <utility>
#include <vector>

static std::mutex g_size_mutex;
static std::unordered_map<const stJpegData*, std::pair<unsigned, unsigned> > g_size_of;

void JpegGetImageSize(stJpegData* jdata, unsigned int* width, unsigned int* height)
{
    unsigned w = 0, h = 0;
    {
        std::lock_guard<std::mutex> lock(g_size_mutex);
        auto it = g_size_of.find(jdata);
        if (it != g_size_of.end()) { w = it->second.first; h = it->second.second; }
    }
    if (width) *width = w;
    if (height) *height = h;
}

static void put16(std::vector<unsigned char>& o, unsigned v) { o.push_back((unsigned char)(v >> 8)); o.push_back((unsigned char)v); }

static void put_dht(std::vector<unsigned char>& o, int cls, int id, const stHuffmanTable& t)
{
    unsigned char bits[16];
    memset(bits, 0, sizeof bits);
    int n = t.m_numBlocks;
    if (n < 0) n = 0;
    if (n > 256) n = 256;
    for (int i = 0; i < n; i++) {
        int len = t.m_blocks[i].length;
        if (len >= 1 && len <= 16) bits[len - 1]++;
    }
    o.push_back(0xFF); o.push_back(0xC4);
    put16(o, (unsigned)(2 + 1 + 16 + n));
    o.push_back((unsigned char)((cls << 4) | id));
    o.insert(o.end(), bits, bits + 16);
    for (int i = 0; i < n; i++) o.push_back((unsigned char)t.m_blocks[i].value);
}

int JpegDecodeHW(stJpegData* jdata, unsigned int jpeg_img_height, unsigned int jpeg_img_width,
                 unsigned char hFactor, unsigned char vFactor)
{
    if (!jdata || !jpeg_img_width || !jpeg_img_height) return 0;
    {
        std::lock_guard<std::mutex> lock(g_size_mutex);
        if (g_size_of.size() > 4096) g_size_of.clear();            // callers that churn through jdata objects
        g_size_of[jdata] = std::make_pair(jpeg_img_width, jpeg_img_height);
    }
    std::vector<unsigned char> o;
    o.reserve(2048 + STREAM_SIZE);
    o.push_back(0xFF); o.push_back(0xD8);
    for (int c = 0; c < 3; c++) {                              // DQT: one table per component
        o.push_back(0xFF); o.push_back(0xDB);
        put16(o, 67);
        o.push_back((unsigned char)c);
        for (int k = 0; k < 64; k++) o.push_back((unsigned char)(int)jdata->m_component_info[cY + c].m_qTable[k]);
    }
    o.push_back(0xFF); o.push_back(0xC0);                      // SOF0
    put16(o, 17);
    o.push_back(8);
    put16(o, jpeg_img_height);
    put16(o, jpeg_img_width);
    o.push_back(3);
    for (int c = 0; c < 3; c++) {
        o.push_back((unsigned char)(cY + c));
        o.push_back(c == 0 ? (unsigned char)((hFactor << 4) | vFactor) : 0x11);
        o.push_back((unsigned char)c);
    }
    bool dc_done[4] = {false, false, false, false}, ac_done[4] = {false, false, false, false};
    for (int c = 0; c < 3; c++) {                              // DHT: the tables the scan selects
        const unsigned d = jdata->m_component_info[cY + c].dcTable_index & 3;
        const unsigned a = jdata->m_component_info[cY + c].acTable_index & 3;
        if (!dc_done[d]) { put_dht(o, 0, (int)d, jdata->m_Huffman.m_HTDC[d]); dc_done[d] = true; }
        if (!ac_done[a]) { put_dht(o, 1, (int)a, jdata->m_Huffman.m_HTAC[a]); ac_done[a] = true; }
    }
    o.push_back(0xFF); o.push_back(0xDA);                      // SOS
    put16(o, 12);
    o.push_back(3);
    for (int c = 0; c < 3; c++) {
        o.push_back((unsigned char)(cY + c));
        o.push_back((unsigned char)(((jdata->m_component_info[cY + c].dcTable_index & 3) << 4) |
                                    (jdata->m_component_info[cY + c].acTable_index & 3)));
    }
    o.push_back(0); o.push_back(63); o.push_back(0);
    o.insert(o.end(), jdata->m_Huffman.m_stream, jdata->m_Huffman.m_stream + (STREAM_SIZE));

    unsigned char* rgb = 0;
    unsigned w = 0, h = 0;
    if (!hjd_decode_jpg_file_data(o.data(), (int)o.size(), &rgb, &w, &h)) return 0;
    size_t bytes = (size_t)w * h * 3;
    if (bytes > sizeof jdata->m_rgb) bytes = sizeof jdata->m_rgb;
    memcpy(jdata->m_rgb, rgb, bytes);
    hjd_free(rgb);
    return 0;                                                  // the reference returns 0 always (loadjpg.cpp:1189)
}
