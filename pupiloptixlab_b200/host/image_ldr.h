// 8-bit image files for bitmap textures: JPEG, BMP and TGA — the formats stb_image adds to PNG on the reference's
// LDR path (framework/util/texture.cpp:106-117, stbi_load).  Own decoders written from the formats' definitions
// (ITU T.81 for JPEG); stb_image is not taken over.  Where a format leaves the arithmetic to the decoder (JPEG's
// inverse DCT, chroma upsampling and YCbCr -> RGB), the fixed-point rules stb_image documents are followed so that
// a texture yields the same texels here as in the reference (tests/test_image_ldr.py checks that against the
// reference's stb_image compiled as a checker where the reference tree is present).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace Pupil::util::ldr {
struct Pixels8 {
    int w = 0, h = 0, channels = 0; // 1 grey, 2 grey + alpha, 3 RGB, 4 RGBA; row 0 = top of the picture
    std::vector<uint8_t> data;      // interleaved
};
// baseline, extended-sequential and progressive Huffman JPEG, 8 bit, 1, 3 or 4 (CMYK / YCCK, returned as RGB) components
bool LoadJpeg(const uint8_t *file, size_t n, Pixels8 &out, std::string &why);
// Windows bitmaps: 1 / 4 / 8 bit palettised, 16 / 24 / 32 bit direct colour (BI_RGB, BI_BITFIELDS), either row order
bool LoadBmp(const uint8_t *file, size_t n, Pixels8 &out, std::string &why);
// Truevision TGA: colour-mapped, true-colour and grey images, raw or run-length encoded, 8 / 15 / 16 / 24 / 32 bit
bool LoadTga(const uint8_t *file, size_t n, Pixels8 &out, std::string &why);
// binary portable pixmaps: P5 (grey) and P6 (RGB), maxval < 65536 (16-bit samples reduced to the byte stb_image's 8-bit API keeps)
bool LoadPnm(const uint8_t *file, size_t n, Pixels8 &out, std::string &why);
bool LooksLikeTga(const uint8_t *file, size_t n); // TGA has no magic number: header plausibility
}// namespace Pupil::util::ldr
