// Image files for bitmap textures, environment maps and saved frames — the role of util::BitmapTexture::Load / Save
// (framework/util/texture.cpp:13-174), which lean on stb_image / stb_image_write / tinyexr.  Those libraries are not
// taken over; this is an own reader/writer for the formats the reference's scenes and output use:
//   read : .hdr (Radiance RGBE, flat + new-style RLE), .exr (scan-line or single-level tiled, NONE / RLE / ZIPS / ZIP / PIZ, HALF or FLOAT channels),
//          .png (1-16 bit, grey / RGB / palette / alpha, Adam7), .jpg (baseline + progressive), .bmp, .tga, .pgm / .ppm (image_ldr.cpp), .pfm
//   write: .hdr (RGBE), .exr (3 FLOAT channels B, G, R, ZIP), .pfm
// Conventions kept from the reference: texels are RGBA float, row 0 = first row of the file (no flip on load); 8-bit
// sources are linearised with pow(x / 255, 2.2) and alpha / 255 (texture.cpp:104-115); HDR sources are taken as they
// are with alpha 1; Save flips vertically because frame buffers have row 0 at the bottom (texture.cpp:14, :37).
#pragma once
#include <cstddef>
#include <string>
#include <string_view>
#include <vector>

namespace Pupil::util {
struct Image {
    size_t w = 0, h = 0;
    std::vector<float> rgba; // w * h * 4
    bool Valid() const noexcept { return w && h && rgba.size() == w * h * 4; }
};

// BitmapTexture::Load: false (and a warning) when the file is missing, malformed or of an unsupported kind
bool LoadImage(std::string_view path, Image &out) noexcept;

enum class EImageFileFormat { HDR, EXR, PFM, PNG };
// BitmapTexture::Save: `data` is w*h RGBA float, row 0 = bottom of the picture.  PNG stores what the reference's canvas shows
// (framework/system/gui/output.hlsl:30-72, gui.cpp:60-61): optional ACES tone mapping, gamma 2.2 (defaults: off, on), 8 bit.
struct DisplayTransform {
    bool tone_mapping = false, gamma_correct = true;
};
bool SaveImage(const float *data, size_t w, size_t h, std::string_view path, EImageFileFormat format, DisplayTransform display = {}) noexcept;
// output.hlsl PSMain for one pixel: rgb in, display-referred rgb in [0, 1] out (not clamped by the shader; the render target is)
void DisplayColor(const float rgb_in[3], DisplayTransform display, float rgb_out[3]) noexcept;
}// namespace Pupil::util
