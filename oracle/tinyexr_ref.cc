// ORACLE — test infrastructure only.  The reference reads .exr textures and environment maps with tinyexr's LoadEXR
// (framework/util/texture.cpp:129-152).  This shim compiles that library from the reference tree where it lies
// (3rdparty/exr/tinyexr/tinyexr.h + deps/miniz, nothing is copied) into oracle/_ref/libtinyexr_ref.so: tests use it to
// WRITE files in every compression the host library claims to read (PIZ has no other encoder in this image) and to hold the
// host library's own EXR reader to what the reference's LoadEXR returns.  Built only where the reference tree exists.
#define TINYEXR_IMPLEMENTATION
#include "tinyexr.h"

#include <cstdlib>
#include <cstring>
#include <vector>

extern "C" {
// rgba: w*h*4 floats out (malloc'ed by tinyexr, free with exr_ref_free); returns tinyexr's status (0 = ok)
int exr_ref_load(const char *path, float **rgba, int *w, int *h) {
    const char *err = nullptr;
    const int rc = LoadEXR(rgba, w, h, path, &err);
    if (err) FreeEXRErrorMessage(err);
    return rc;
}
void exr_ref_free(void *p) { std::free(p); }
// writes w*h pixels of `channels` (1 = Y, 3 = RGB, 4 = RGBA) interleaved floats as a scan-line file with the given
// TINYEXR_COMPRESSIONTYPE_* and HALF or FLOAT channels
int exr_ref_save(const char *path, const float *data, int w, int h, int channels, int compression, int half) {
    EXRHeader header;
    InitEXRHeader(&header);
    EXRImage image;
    InitEXRImage(&image);
    const size_t n = static_cast<size_t>(w) * h;
    std::vector<std::vector<float>> planes(channels, std::vector<float>(n));
    for (size_t i = 0; i < n; ++i)
        for (int c = 0; c < channels; ++c) planes[c][i] = data[i * channels + c];
    // channels must be stored in alphabetical order: A, B, G, R (or Y)
    static const char *names4[4] = { "A", "B", "G", "R" }, *names3[3] = { "B", "G", "R" }, *names1[1] = { "Y" };
    static const int src4[4] = { 3, 2, 1, 0 }, src3[3] = { 2, 1, 0 }, src1[1] = { 0 };
    const char **names = channels == 4 ? names4 : channels == 3 ? names3 : names1;
    const int *src = channels == 4 ? src4 : channels == 3 ? src3 : src1;
    std::vector<unsigned char *> ptrs(channels);
    std::vector<EXRChannelInfo> infos(channels);
    std::vector<int> in_types(channels, TINYEXR_PIXELTYPE_FLOAT), out_types(channels, half ? TINYEXR_PIXELTYPE_HALF : TINYEXR_PIXELTYPE_FLOAT);
    for (int c = 0; c < channels; ++c) {
        ptrs[c] = reinterpret_cast<unsigned char *>(planes[src[c]].data());
        std::memset(&infos[c], 0, sizeof(EXRChannelInfo));
        std::strncpy(infos[c].name, names[c], 255);
    }
    image.images = ptrs.data(), image.width = w, image.height = h, image.num_channels = channels;
    header.num_channels = channels, header.channels = infos.data();
    header.pixel_types = in_types.data(), header.requested_pixel_types = out_types.data();
    header.compression_type = compression;
    const char *err = nullptr;
    const int rc = SaveEXRImageToFile(&image, &header, path, &err);
    if (err) FreeEXRErrorMessage(err);
    return rc;
}
}
