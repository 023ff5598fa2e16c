// PIZ block decompression for the OpenEXR reader (image.cpp); see image_piz.cpp
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace Pupil::util::piz {
// one scan-line block (up to 32 lines of `width` pixels); words_per_sample: 1 for a HALF channel, 2 for FLOAT / UINT, in file
// (alphabetical) channel order.  out: the block's bytes in EXR scan-line layout (per line, channel after channel, little endian)
bool Decompress(const uint8_t *in, size_t n_in, size_t width, size_t lines, const std::vector<int> &words_per_sample, std::vector<uint8_t> &out);
}// namespace Pupil::util::piz
