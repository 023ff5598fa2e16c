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
// the same picture as a single-level tiled file (tile_w x tile_h tiles)
int exr_ref_save_tiled(const char *path, const float *data, int w, int h, int channels, int compression, int half, int tile_w, int tile_h) {
    EXRHeader header;
    InitEXRHeader(&header);
    EXRImage image;
    InitEXRImage(&image);
    static const char *names4[4] = { "A", "B", "G", "R" }, *names3[3] = { "B", "G", "R" }, *names1[1] = { "Y" };
    static const int src4[4] = { 3, 2, 1, 0 }, src3[3] = { 2, 1, 0 }, src1[1] = { 0 };
    const char **names = channels == 4 ? names4 : channels == 3 ? names3 : names1;
    const int *src = channels == 4 ? src4 : channels == 3 ? src3 : src1;
    const int nx = (w + tile_w - 1) / tile_w, ny = (h + tile_h - 1) / tile_h;
    std::vector<EXRTile> tiles(static_cast<size_t>(nx) * ny);
    std::vector<std::vector<float>> store(tiles.size() * channels, std::vector<float>(static_cast<size_t>(tile_w) * tile_h, 0.f));
    std::vector<std::vector<unsigned char *>> ptrs(tiles.size(), std::vector<unsigned char *>(channels));
    for (int ty = 0; ty < ny; ++ty)
        for (int tx = 0; tx < nx; ++tx) {
            const size_t t = static_cast<size_t>(ty) * nx + tx;
            EXRTile &tile = tiles[t];
            tile.offset_x = tx, tile.offset_y = ty, tile.level_x = 0, tile.level_y = 0;
            tile.width = std::min(tile_w, w - tx * tile_w), tile.height = std::min(tile_h, h - ty * tile_h);
            for (int c = 0; c < channels; ++c) {
                std::vector<float> &plane = store[t * channels + c];
                for (int y = 0; y < tile.height; ++y)
                    for (int x = 0; x < tile.width; ++x)
                        plane[static_cast<size_t>(y) * tile_w + x] = data[(static_cast<size_t>(ty * tile_h + y) * w + tx * tile_w + x) * channels + src[c]];
                ptrs[t][c] = reinterpret_cast<unsigned char *>(plane.data());
            }
            tile.images = ptrs[t].data();
        }
    std::vector<EXRChannelInfo> infos(channels);
    std::vector<int> in_types(channels, TINYEXR_PIXELTYPE_FLOAT), out_types(channels, half ? TINYEXR_PIXELTYPE_HALF : TINYEXR_PIXELTYPE_FLOAT);
    for (int c = 0; c < channels; ++c) {
        std::memset(&infos[c], 0, sizeof(EXRChannelInfo));
        std::strncpy(infos[c].name, names[c], 255);
    }
    image.tiles = tiles.data(), image.num_tiles = static_cast<int>(tiles.size()), image.images = nullptr;
    image.width = w, image.height = h, image.num_channels = channels, image.level_x = 0, image.level_y = 0;
    header.num_channels = channels, header.channels = infos.data();
    header.pixel_types = in_types.data(), header.requested_pixel_types = out_types.data();
    header.compression_type = compression;
    header.data_window.min_x = header.data_window.min_y = header.display_window.min_x = header.display_window.min_y = 0;
    header.data_window.max_x = header.display_window.max_x = w - 1, header.data_window.max_y = header.display_window.max_y = h - 1;
    header.tiled = 1, header.tile_size_x = tile_w, header.tile_size_y = tile_h;
    header.tile_level_mode = TINYEXR_TILE_ONE_LEVEL, header.tile_rounding_mode = TINYEXR_TILE_ROUND_DOWN;
    const char *err = nullptr;
    const int rc = SaveEXRImageToFile(&image, &header, path, &err);
    if (err) FreeEXRErrorMessage(err);
    return rc;
}
}
