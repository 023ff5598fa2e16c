// PIZ decompression for the OpenEXR reader in image.cpp — the compression most tools write .exr environment maps with, and
// one tinyexr's LoadEXR (the reference's reader, framework/util/texture.cpp:131-149) handles.  Written from the published
// description of the scheme (OpenEXR technical introduction; Imf PIZ: a 16-bit value range compaction through a bitmap-derived
// lookup table, a two-dimensional Haar-like wavelet per channel plane, canonical Huffman coding with zero-run and repeat
// symbols); tinyexr's code is not taken over.  tests/test_image_exr.py holds the result to tinyexr compiled from the
// reference tree as a checker, which also writes the test files.
#include "image_piz.h"

#include <cstring>

namespace Pupil::util::piz {
namespace {
constexpr int kEncBits = 16, kDecBits = 14;
constexpr int kEncSize = (1 << kEncBits) + 1; // symbols 0..65535 plus the run-length symbol
constexpr int kDecSize = 1 << kDecBits, kDecMask = kDecSize - 1;
constexpr int kShortZeroRun = 59, kLongZeroRun = 63, kShortestLongRun = 2 + kLongZeroRun - kShortZeroRun;
using U64 = unsigned long long;

uint32_t Le32(const uint8_t *p) { return p[0] | p[1] << 8 | p[2] << 16 | static_cast<uint32_t>(p[3]) << 24; }

struct DecEntry {
    uint32_t len = 0;           // code length of a short code (<= 14 bits), 0 for a long-code bucket
    uint32_t lit = 0;           // the symbol, or the number of long codes sharing this 14-bit prefix
    std::vector<uint32_t> list; // those long codes' symbols
};

// code lengths, 6 bits each, with run-length escapes for zero lengths; then the canonical code of every symbol in (code << 6 | length)
bool UnpackEncTable(const uint8_t *&p, const uint8_t *end, uint32_t im, uint32_t iM, std::vector<U64> &hcode) {
    U64 c = 0;
    int lc = 0;
    auto bits = [&](int n, U64 &out) -> bool {
        while (lc < n) {
            if (p >= end) return false;
            c = (c << 8) | *p++, lc += 8;
        }
        lc -= n;
        out = (c >> lc) & ((1ull << n) - 1);
        return true;
    };
    for (; im <= iM; ++im) {
        U64 l;
        if (!bits(6, l)) return false;
        hcode[im] = l;
        U64 run = 0;
        if (l == static_cast<U64>(kLongZeroRun)) {
            if (!bits(8, run)) return false;
            run += kShortestLongRun;
        } else if (l >= static_cast<U64>(kShortZeroRun)) {
            run = l - kShortZeroRun + 2;
        }
        if (run) {
            if (im + run > static_cast<U64>(iM) + 1) return false;
            while (run--) hcode[im++] = 0;
            --im;
        }
    }
    // canonical codes: within a length in symbol order, longer codes numerically first
    U64 n[59] = {};
    for (int i = 0; i < kEncSize; ++i) {
        if (hcode[i] > 58) return false;
        ++n[hcode[i]];
    }
    U64 code = 0;
    for (int i = 58; i > 0; --i) {
        const U64 next = (code + n[i]) >> 1;
        n[i] = code, code = next;
    }
    for (int i = 0; i < kEncSize; ++i) {
        const U64 l = hcode[i];
        if (l > 0) hcode[i] = l | (n[l]++ << 6);
    }
    return true;
}

bool BuildDecTable(const std::vector<U64> &hcode, uint32_t im, uint32_t iM, std::vector<DecEntry> &dec) {
    for (; im <= iM; ++im) {
        const U64 c = hcode[im] >> 6;
        const int l = static_cast<int>(hcode[im] & 63);
        if (c >> l) return false;
        if (l > kDecBits) {
            DecEntry &e = dec[c >> (l - kDecBits)];
            if (e.len) return false;
            ++e.lit, e.list.push_back(im);
        } else if (l) {
            DecEntry *e = &dec[c << (kDecBits - l)];
            for (U64 i = 1ull << (kDecBits - l); i > 0; --i, ++e) {
                if (e->len || !e->list.empty()) return false;
                e->len = static_cast<uint32_t>(l), e->lit = im;
            }
        }
    }
    return true;
}

bool HufDecode(const std::vector<U64> &hcode, const std::vector<DecEntry> &dec, const uint8_t *in, U64 n_bits, uint32_t rlc, size_t n_out, uint16_t *out) {
    U64 c = 0;
    int lc = 0;
    uint16_t *const begin = out, *const oe = out + n_out;
    const uint8_t *const ie = in + (n_bits + 7) / 8;
    auto emit = [&](uint32_t sym) -> bool {
        if (sym == rlc) { // repeat the previous value: count in the next 8 bits
            if (lc < 8) {
                if (in >= ie) return false;
                c = (c << 8) | *in++, lc += 8;
            }
            lc -= 8;
            size_t run = (c >> lc) & 0xff;
            if (out + run > oe || out == begin) return false;
            const uint16_t s = out[-1];
            while (run--) *out++ = s;
        } else {
            if (out >= oe) return false;
            *out++ = static_cast<uint16_t>(sym);
        }
        return true;
    };
    while (in < ie) {
        c = (c << 8) | *in++, lc += 8;
        while (lc >= kDecBits) {
            const DecEntry &e = dec[(c >> (lc - kDecBits)) & kDecMask];
            if (e.len) {
                lc -= static_cast<int>(e.len);
                if (!emit(e.lit)) return false;
            } else {
                if (e.list.empty()) return false;
                size_t j = 0;
                for (; j < e.list.size(); ++j) {
                    const int l = static_cast<int>(hcode[e.list[j]] & 63);
                    while (lc < l && in < ie) c = (c << 8) | *in++, lc += 8;
                    if (lc >= l && (hcode[e.list[j]] >> 6) == ((c >> (lc - l)) & ((1ull << l) - 1))) {
                        lc -= l;
                        if (!emit(e.list[j])) return false;
                        break;
                    }
                }
                if (j == e.list.size()) return false;
            }
        }
    }
    const int pad = static_cast<int>((8 - n_bits) & 7); // bits of the last byte that are not data
    c >>= pad, lc -= pad;
    while (lc > 0) {
        const DecEntry &e = dec[(c << (kDecBits - lc)) & kDecMask];
        if (!e.len) return false;
        lc -= static_cast<int>(e.len);
        if (lc < 0 || !emit(e.lit)) return false;
    }
    return out == oe;
}

bool HufUncompress(const uint8_t *data, size_t n, uint16_t *raw, size_t n_raw) {
    if (n == 0) return n_raw == 0;
    if (n < 20) return false;
    const uint32_t im = Le32(data), iM = Le32(data + 4), n_bits = Le32(data + 12);
    if (im >= static_cast<uint32_t>(kEncSize) || iM >= static_cast<uint32_t>(kEncSize) || im > iM) return false;
    const uint8_t *p = data + 20, *end = data + n;
    std::vector<U64> hcode(kEncSize, 0);
    if (!UnpackEncTable(p, end, im, iM, hcode)) return false;
    if (n_bits > 8 * static_cast<U64>(end - p)) return false;
    std::vector<DecEntry> dec(kDecSize);
    if (!BuildDecTable(hcode, im, iM, dec)) return false;
    return HufDecode(hcode, dec, p, n_bits, iM, n_raw, raw);
}

// inverse of the two-point transforms: (average, difference) -> (a, b); 14-bit data in plain integer arithmetic, 16-bit data modulo 2^16
inline void Dec14(uint16_t l, uint16_t h, uint16_t &a, uint16_t &b) {
    const int ls = static_cast<int16_t>(l), hs = static_cast<int16_t>(h);
    const int ai = ls + (hs & 1) + (hs >> 1);
    a = static_cast<uint16_t>(static_cast<int16_t>(ai)), b = static_cast<uint16_t>(static_cast<int16_t>(ai - hs));
}
inline void Dec16(uint16_t l, uint16_t h, uint16_t &a, uint16_t &b) {
    const int m = l, d = h;
    const int bb = (m - (d >> 1)) & 0xffff, aa = (d + bb - 0x8000) & 0xffff;
    b = static_cast<uint16_t>(bb), a = static_cast<uint16_t>(aa);
}
// in-place inverse wavelet of an nx x ny plane whose samples are ox apart in x and oy apart in y
void WaveletDecode(uint16_t *in, int nx, int ox, int ny, int oy, uint16_t max_value) {
    const bool w14 = max_value < (1 << 14);
    const int n = nx > ny ? ny : nx;
    int p = 1;
    while (p <= n) p <<= 1;
    p >>= 1;
    int p2 = p;
    p >>= 1;
    auto dec = [&](uint16_t l, uint16_t h, uint16_t &a, uint16_t &b) { w14 ? Dec14(l, h, a, b) : Dec16(l, h, a, b); };
    while (p >= 1) {
        uint16_t *py = in, *const ey = in + static_cast<ptrdiff_t>(oy) * (ny - p2);
        const ptrdiff_t oy1 = static_cast<ptrdiff_t>(oy) * p, oy2 = static_cast<ptrdiff_t>(oy) * p2, ox1 = static_cast<ptrdiff_t>(ox) * p, ox2 = static_cast<ptrdiff_t>(ox) * p2;
        for (; py <= ey; py += oy2) {
            uint16_t *px = py, *const ex = py + static_cast<ptrdiff_t>(ox) * (nx - p2);
            for (; px <= ex; px += ox2) {
                uint16_t *p01 = px + ox1, *p10 = px + oy1, *p11 = p10 + ox1;
                uint16_t i00, i01, i10, i11;
                dec(*px, *p10, i00, i10), dec(*p01, *p11, i01, i11);
                dec(i00, i01, *px, *p01), dec(i10, i11, *p10, *p11);
            }
            if (nx & p) { // an odd column is left: transform in y only
                uint16_t *p10 = px + oy1, i00;
                dec(*px, *p10, i00, *p10);
                *px = i00;
            }
        }
        if (ny & p) { // an odd row is left: transform in x only
            uint16_t *px = py, *const ex = py + static_cast<ptrdiff_t>(ox) * (nx - p2);
            for (; px <= ex; px += ox2) {
                uint16_t *p01 = px + ox1, i00;
                dec(*px, *p01, i00, *p01);
                *px = i00;
            }
        }
        p2 = p, p >>= 1;
    }
}
}// namespace

bool Decompress(const uint8_t *in, size_t n_in, size_t width, size_t lines, const std::vector<int> &words_per_sample, std::vector<uint8_t> &out) {
    size_t words_per_line = 0;
    for (int w : words_per_sample) words_per_line += static_cast<size_t>(w) * width;
    const size_t n_words = words_per_line * lines;
    if (n_in < 4) return false;
    std::vector<uint8_t> bitmap(8192, 0);
    const uint32_t min_nz = in[0] | in[1] << 8, max_nz = in[2] | in[3] << 8;
    const uint8_t *p = in + 4, *const end = in + n_in;
    if (max_nz >= 8192) return false;
    if (min_nz <= max_nz) {
        const size_t len = max_nz - min_nz + 1;
        if (static_cast<size_t>(end - p) < len) return false;
        std::memcpy(bitmap.data() + min_nz, p, len);
        p += len;
    }
    // the values that occur, in ascending order: coded value k stands for lut[k] (zero always occurs)
    std::vector<uint16_t> lut(65536, 0);
    uint32_t k = 0;
    for (uint32_t i = 0; i < 65536; ++i)
        if (i == 0 || (bitmap[i >> 3] & (1u << (i & 7)))) lut[k++] = static_cast<uint16_t>(i);
    const uint16_t max_value = static_cast<uint16_t>(k - 1);
    if (end - p < 4) return false;
    const uint32_t length = Le32(p);
    p += 4;
    if (static_cast<size_t>(end - p) < length) return false;
    std::vector<uint16_t> tmp(n_words);
    if (!HufUncompress(p, length, tmp.data(), n_words)) return false;
    // channel planes follow each other; a 32-bit sample is two interleaved 16-bit planes
    uint16_t *plane = tmp.data();
    std::vector<uint16_t *> cursor;
    for (int w : words_per_sample) {
        cursor.push_back(plane);
        for (int j = 0; j < w; ++j) WaveletDecode(plane + j, static_cast<int>(width), w, static_cast<int>(lines), static_cast<int>(width) * w, max_value);
        plane += width * lines * w;
    }
    for (auto &v : tmp) v = lut[v];
    // back to scan-line order: per line, channel after channel
    out.resize(n_words * 2);
    uint8_t *o = out.data();
    for (size_t y = 0; y < lines; ++y)
        for (size_t c = 0; c < words_per_sample.size(); ++c) {
            const size_t n = width * words_per_sample[c];
            std::memcpy(o, cursor[c], n * 2);
            o += n * 2, cursor[c] += n;
        }
    return true;
}
}// namespace Pupil::util::piz
