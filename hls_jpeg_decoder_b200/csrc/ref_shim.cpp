// csrc/ref_shim.cpp -- the reference's own C++ entry points, defined on top of the C ABI.
//
// harutel/hls-jpeg-decoder has plain C++ linkage and no namespaces, so a program
// written against it (its src/main.cpp, or anything calling ConvertJpgFile) links against
// libhjd.so unchanged: same names, same argument meaning, same return convention
// (1 = success, 0 = failure; openjpg.cpp:607,643,665,683).
//
//   ConvertJpgFile     openjpg.h:23 / openjpg.cpp:593
//   DecodeJpgFileData  loadjpg.h:186 / openjpg.h:26 -- declared by the reference, never defined;
//                      implemented here for real.  The caller delete[]s *rgbpix (loadjpg.h:187).
//   WriteBMP24         openjpg.cpp:504 (and the int overload promised at openjpg.h:19)
//
// JpegDecodeHW(stJpegData*, ...) (loadjpg.h:180) takes the reference's fixed-capacity struct and
// therefore needs the reference's header; that adapter lives in ref_shim_hw.cpp and is built
// only where the header is available.
#include "../../include/hjd.h"
#include <string.h>

#if defined(__GNUC__)
#define HJD_EXPORT __attribute__((visibility("default")))
#else
#define HJD_EXPORT
#endif

HJD_EXPORT int ConvertJpgFile(char* szJpgFileInName, char* szBmpFileOutName)
{
    return hjd_convert_jpg_file(szJpgFileInName, szBmpFileOutName);
}

static void* shim_new(size_t bytes) { return new unsigned char[bytes]; }     // "Don't forget to delete[] rgbpix"

HJD_EXPORT int DecodeJpgFileData(const unsigned char* buf, int sizeBuf, unsigned char** rgbpix,
                                 unsigned int* width, unsigned int* height)
{
    if (!rgbpix) return 0;
    // the library copies the result straight into memory from new[]: no intermediate buffer
    return hjd_decode_jpg_file_data_alloc(buf, sizeBuf, shim_new, rgbpix, width, height);
}

HJD_EXPORT void WriteBMP24(const char* szBmpFileName, unsigned int Width, unsigned int Height, unsigned char* RGB)
{
    hjd_write_bmp24(szBmpFileName, Width, Height, RGB);
}

HJD_EXPORT void WriteBMP24(const char* szBmpFileName, int Width, int Height, unsigned char* RGB)
{
    hjd_write_bmp24(szBmpFileName, (unsigned)Width, (unsigned)Height, RGB);
}
