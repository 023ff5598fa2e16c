#include "image.h"
#include "image_ldr.h"
#include "image_piz.h"
#include "util.h"

#include <zlib.h>

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <filesystem>

namespace Pupil::util {
namespace {
bool ReadFile(std::string_view path, std::vector<uint8_t> &out) {
    std::FILE *f = std::fopen(std::string(path).c_str(), "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    bool ok = n >= 0;
    if (ok) {
        out.resize(static_cast<size_t>(n));
        ok = n == 0 || std::fread(out.data(), 1, out.size(), f) == out.size();
    }
    std::fclose(f);
    return ok;
}
struct Reader {
    const uint8_t *p, *end;
    bool ok = true;
    Reader(const uint8_t *b, size_t n) : p(b), end(b + n) {}
    size_t left() const { return static_cast<size_t>(end - p); }
    const uint8_t *take(size_t n) {
        if (left() < n) {
            ok = false;
            return nullptr;
        }
        const uint8_t *r = p;
        p += n;
        return r;
    }
    template<typename T>
    T le() { // little-endian scalar
        T v{};
        if (const uint8_t *q = take(sizeof(T))) std::memcpy(&v, q, sizeof(T));
        return v;
    }
    uint32_t be32() {
        const uint8_t *q = take(4);
        return q ? (uint32_t(q[0]) << 24 | uint32_t(q[1]) << 16 | uint32_t(q[2]) << 8 | q[3]) : 0u;
    }
    std::string cstr() { // zero-terminated
        std::string s;
        while (p < end && *p) s.push_back(static_cast<char>(*p++));
        if (p < end) ++p;
        else ok = false;
        return s;
    }
    std::string line() { // up to and excluding '\n'
        std::string s;
        while (p < end && *p != '\n') s.push_back(static_cast<char>(*p++));
        if (p < end) ++p;
        return s;
    }
};
float LdrToLinear(uint8_t v) { return std::pow(v * 1.f / 255.f, 2.2f); } // texture.cpp:108-110
// 8-bit pixels -> RGBA float the way StbImageLoad does (texture.cpp:106-117): colour through pow(x / 255, 2.2), alpha / 255,
// alpha 1 without an alpha channel; grey is spread over R, G, B
void FromPixels8(const ldr::Pixels8 &px, Image &img) {
    float lut[256];
    for (int i = 0; i < 256; ++i) lut[i] = LdrToLinear(static_cast<uint8_t>(i));
    img.w = static_cast<size_t>(px.w), img.h = static_cast<size_t>(px.h);
    img.rgba.resize(img.w * img.h * 4);
    const int c = px.channels;
    for (size_t i = 0, n = img.w * img.h; i < n; ++i) {
        const uint8_t *s = px.data.data() + i * c;
        float *o = &img.rgba[i * 4];
        if (c >= 3) o[0] = lut[s[0]], o[1] = lut[s[1]], o[2] = lut[s[2]], o[3] = c == 4 ? s[3] * 1.f / 255.f : 1.f;
        else o[0] = o[1] = o[2] = lut[s[0]], o[3] = c == 2 ? s[1] * 1.f / 255.f : 1.f;
    }
}

// ---- PFM ("PF" colour / "Pf" grey, bottom-to-top rows, negative scale = little endian) ---------------------------
bool LoadPfm(const std::vector<uint8_t> &file, Image &img) {
    Reader r(file.data(), file.size());
    const std::string magic = r.line();
    if (magic.rfind("PF", 0) != 0 && magic.rfind("Pf", 0) != 0) return false;
    const int channels = magic[1] == 'F' ? 3 : 1;
    int w = 0, h = 0;
    float scale = 0.f;
    const std::string dims = r.line();
    if (std::sscanf(dims.c_str(), "%d %d", &w, &h) != 2 || w <= 0 || h <= 0) return false;
    if (std::sscanf(r.line().c_str(), "%f", &scale) != 1 || scale == 0.f) return false;
    const size_t n = static_cast<size_t>(w) * h * channels;
    const uint8_t *data = r.take(n * 4);
    if (!data) return false;
    img.w = w, img.h = h;
    img.rgba.resize(static_cast<size_t>(w) * h * 4);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float c[3];
            for (int k = 0; k < channels; ++k) {
                uint8_t b[4];
                std::memcpy(b, data + ((static_cast<size_t>(h - 1 - y) * w + x) * channels + k) * 4, 4); // file rows run bottom-up
                if (scale > 0.f) std::swap(b[0], b[3]), std::swap(b[1], b[2]);
                std::memcpy(&c[k], b, 4);
            }
            float *o = &img.rgba[(static_cast<size_t>(y) * w + x) * 4];
            o[0] = c[0], o[1] = channels == 3 ? c[1] : c[0], o[2] = channels == 3 ? c[2] : c[0], o[3] = 1.f;
        }
    return true;
}

// ---- Radiance .hdr -----------------------------------------------------------------------------------------------
bool LoadHdr(const std::vector<uint8_t> &file, Image &img) {
    Reader r(file.data(), file.size());
    const std::string magic = r.line();
    if (magic != "#?RADIANCE" && magic != "#?RGBE") return false;
    bool format_ok = false;
    for (;;) {
        if (r.left() == 0) return false;
        const std::string l = r.line();
        if (l.empty()) break;
        if (l == "FORMAT=32-bit_rle_rgbe") format_ok = true;
    }
    if (!format_ok) return false;
    int w = 0, h = 0;
    if (std::sscanf(r.line().c_str(), "-Y %d +X %d", &h, &w) != 2 || w <= 0 || h <= 0) return false;
    img.w = w, img.h = h;
    img.rgba.resize(static_cast<size_t>(w) * h * 4);
    std::vector<uint8_t> scan(static_cast<size_t>(w) * 4);
    auto convert = [&](int y) {
        for (int x = 0; x < w; ++x) {
            const uint8_t *e = &scan[static_cast<size_t>(x) * 4];
            float *o = &img.rgba[(static_cast<size_t>(y) * w + x) * 4];
            if (e[3]) {
                const float f = std::ldexp(1.0f, static_cast<int>(e[3]) - (128 + 8));
                o[0] = e[0] * f, o[1] = e[1] * f, o[2] = e[2] * f;
            } else {
                o[0] = o[1] = o[2] = 0.f;
            }
            o[3] = 1.f;
        }
    };
    for (int y = 0; y < h; ++y) {
        bool rle = false;
        if (w >= 8 && w < 32768 && r.left() >= 4 && r.p[0] == 2 && r.p[1] == 2 && !(r.p[2] & 0x80)) {
            rle = (static_cast<int>(r.p[2]) << 8 | r.p[3]) == w;
        }
        if (!rle) { // flat scanline
            const uint8_t *q = r.take(static_cast<size_t>(w) * 4);
            if (!q) return false;
            std::memcpy(scan.data(), q, static_cast<size_t>(w) * 4);
        } else {
            r.take(4);
            for (int k = 0; k < 4; ++k) {
                int x = 0;
                while (x < w) {
                    const uint8_t *c = r.take(1);
                    if (!c) return false;
                    int count = *c;
                    if (count > 128) { // run
                        count -= 128;
                        const uint8_t *v = r.take(1);
                        if (!v || x + count > w) return false;
                        for (int i = 0; i < count; ++i) scan[static_cast<size_t>(x++) * 4 + k] = *v;
                    } else { // literal
                        const uint8_t *v = r.take(count);
                        if (!v || count == 0 || x + count > w) return false;
                        for (int i = 0; i < count; ++i) scan[static_cast<size_t>(x++) * 4 + k] = v[i];
                    }
                }
            }
        }
        convert(y);
    }
    return true;
}
void FloatToRgbe(const float *c, uint8_t *out) {
    const float m = std::max(c[0], std::max(c[1], c[2]));
    if (m < 1e-32f) {
        out[0] = out[1] = out[2] = out[3] = 0;
        return;
    }
    int e;
    const float norm = std::frexp(m, &e) * 256.0f / m;
    out[0] = static_cast<uint8_t>(c[0] * norm), out[1] = static_cast<uint8_t>(c[1] * norm), out[2] = static_cast<uint8_t>(c[2] * norm);
    out[3] = static_cast<uint8_t>(e + 128);
}
bool SaveHdr(const float *data, size_t w, size_t h, const std::string &path) {
    // C stdio on purpose: formatted iostream output crashed inside host processes that had loaded another C++ runtime first
    std::FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    bool ok = std::fprintf(f, "#?RADIANCE\n# Written by pupiloptixlab_b200\nFORMAT=32-bit_rle_rgbe\nEXPOSURE=1.0\n\n-Y %zu +X %zu\n", h, w) > 0;
    std::vector<uint8_t> scan(w * 4);
    for (size_t y = 0; y < h && ok; ++y) {
        const float *row = data + (h - 1 - y) * w * 4; // flip: buffer row 0 is the bottom of the picture
        for (size_t x = 0; x < w; ++x) FloatToRgbe(row + x * 4, &scan[x * 4]);
        ok = std::fwrite(scan.data(), 1, scan.size(), f) == scan.size();
    }
    return std::fclose(f) == 0 && ok;
}
bool SavePfm(const float *data, size_t w, size_t h, const std::string &path) {
    std::FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    bool ok = std::fprintf(f, "PF\n%zu %zu\n-1.0\n", w, h) > 0;
    std::vector<float> row(w * 3);
    for (size_t y = 0; y < h && ok; ++y) { // PFM rows run bottom-up, like the buffer
        for (size_t x = 0; x < w; ++x)
            for (int k = 0; k < 3; ++k) row[x * 3 + k] = data[(y * w + x) * 4 + k];
        ok = std::fwrite(row.data(), 4, row.size(), f) == row.size();
    }
    return std::fclose(f) == 0 && ok;
}

// ---- PNG (ISO/IEC 15948): zlib stream of filtered scanlines ---------------------------------------------------------
uint8_t Paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return static_cast<uint8_t>(pa <= pb && pa <= pc ? a : (pb <= pc ? b : c));
}
bool LoadPng(const std::vector<uint8_t> &file, Image &img) {
    static const uint8_t sig[8] = { 0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a };
    if (file.size() < 8 || std::memcmp(file.data(), sig, 8) != 0) return false;
    Reader r(file.data() + 8, file.size() - 8);
    uint32_t w = 0, h = 0;
    int depth = 0, color = 0, interlace = 0;
    std::vector<uint8_t> idat, palette, trns;
    for (;;) {
        const uint32_t len = r.be32();
        const uint8_t *type = r.take(4);
        if (!r.ok || !type) return false;
        const uint8_t *body = r.take(len);
        r.take(4); // CRC (not verified, like stb_image)
        if (!r.ok) return false;
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len != 13) return false;
            w = uint32_t(body[0]) << 24 | uint32_t(body[1]) << 16 | uint32_t(body[2]) << 8 | body[3];
            h = uint32_t(body[4]) << 24 | uint32_t(body[5]) << 16 | uint32_t(body[6]) << 8 | body[7];
            depth = body[8], color = body[9], interlace = body[12];
        } else if (!std::memcmp(type, "PLTE", 4)) {
            palette.assign(body, body + len);
        } else if (!std::memcmp(type, "tRNS", 4)) {
            trns.assign(body, body + len);
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            break;
        }
    }
    if (!w || !h || interlace > 1) return false;
    int channels;
    switch (color) {
        case 0: channels = 1; break;
        case 2: channels = 3; break;
        case 3: channels = 1; break;
        case 4: channels = 2; break;
        case 6: channels = 4; break;
        default: return false;
    }
    if (!(depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) return false;
    if (color == 3 && (depth == 16 || palette.empty())) return false;
    const size_t bpp_bits = static_cast<size_t>(channels) * depth, bpp = std::max<size_t>(1, bpp_bits / 8);
    auto stride_of = [&](uint32_t pw) { return (pw * bpp_bits + 7) / 8; };
    // the seven Adam7 passes (or the one pass of a non-interlaced file): origin and step of the sub-image inside the picture
    struct Pass {
        uint32_t x0, y0, dx, dy;
    };
    static const Pass adam7[7] = { { 0, 0, 8, 8 }, { 4, 0, 8, 8 }, { 0, 4, 4, 8 }, { 2, 0, 4, 4 }, { 0, 2, 2, 4 }, { 1, 0, 2, 2 }, { 0, 1, 1, 2 } };
    static const Pass whole = { 0, 0, 1, 1 };
    const Pass *passes = interlace ? adam7 : &whole;
    const int n_passes = interlace ? 7 : 1;
    auto pass_w = [&](const Pass &ps) { return ps.x0 < w ? (w - ps.x0 + ps.dx - 1) / ps.dx : 0u; };
    auto pass_h = [&](const Pass &ps) { return ps.y0 < h ? (h - ps.y0 + ps.dy - 1) / ps.dy : 0u; };
    size_t raw_size = 0;
    for (int i = 0; i < n_passes; ++i)
        if (pass_w(passes[i]) && pass_h(passes[i])) raw_size += pass_h(passes[i]) * (stride_of(pass_w(passes[i])) + 1);
    // deflate expands by at most ~1032 : 1, so a header that promises more than the IDAT bytes can hold is corrupt (and must not
    // be able to ask for gigabytes)
    if (static_cast<uint64_t>(w) * h > (1ull << 28) || raw_size > idat.size() * 1040 + 64) return false;
    std::vector<uint8_t> raw(raw_size);
    uLongf raw_len = static_cast<uLongf>(raw.size());
    if (uncompress(raw.data(), &raw_len, idat.data(), static_cast<uLong>(idat.size())) != Z_OK || raw_len != raw.size()) return false;
    img.w = w, img.h = h;
    img.rgba.resize(static_cast<size_t>(w) * h * 4);
    const uint8_t *src = raw.data();
    for (int pi = 0; pi < n_passes; ++pi) {
        const Pass &ps = passes[pi];
        const uint32_t pw = pass_w(ps), ph = pass_h(ps);
        if (!pw || !ph) continue; // an empty pass has no bytes at all, not even filter bytes
        const size_t stride = stride_of(pw);
        std::vector<uint8_t> prev(stride, 0), cur(stride);
        for (uint32_t y = 0; y < ph; ++y) {
            const uint8_t *line = src + y * (stride + 1);
            const int filter = line[0];
            for (size_t i = 0; i < stride; ++i) {
                const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
                int v = line[1 + i];
                switch (filter) {
                    case 0: break;
                    case 1: v += a; break;
                    case 2: v += b; break;
                    case 3: v += (a + b) >> 1; break;
                    case 4: v += Paeth(a, b, c); break;
                    default: return false;
                }
                cur[i] = static_cast<uint8_t>(v);
            }
            auto sample8 = [&](uint32_t x, int ch) -> uint8_t { // sample reduced to 8 bits the way stb_image's 8-bit API does
                if (depth == 8) return cur[static_cast<size_t>(x) * channels + ch];
                if (depth == 16) return cur[(static_cast<size_t>(x) * channels + ch) * 2]; // high byte
                const size_t bit = static_cast<size_t>(x) * depth;
                const uint8_t v = (cur[bit / 8] >> (8 - depth - bit % 8)) & ((1u << depth) - 1u);
                return color == 3 ? v : static_cast<uint8_t>(v * (255 / ((1 << depth) - 1)));
            };
            for (uint32_t x = 0; x < pw; ++x) {
                uint8_t px[4] = { 0, 0, 0, 255 };
                if (color == 3) {
                    const uint8_t idx = sample8(x, 0);
                    if (static_cast<size_t>(idx) * 3 + 2 < palette.size()) px[0] = palette[idx * 3], px[1] = palette[idx * 3 + 1], px[2] = palette[idx * 3 + 2];
                    if (idx < trns.size()) px[3] = trns[idx];
                } else if (channels <= 2) { // grey (+alpha): expanded to RGB (the reference indexes 3 channels regardless)
                    px[0] = px[1] = px[2] = sample8(x, 0);
                    if (channels == 2) px[3] = sample8(x, 1);
                } else {
                    px[0] = sample8(x, 0), px[1] = sample8(x, 1), px[2] = sample8(x, 2);
                    if (channels == 4) px[3] = sample8(x, 3);
                }
                float *o = &img.rgba[(static_cast<size_t>(ps.y0 + y * ps.dy) * w + ps.x0 + x * ps.dx) * 4];
                o[0] = LdrToLinear(px[0]), o[1] = LdrToLinear(px[1]), o[2] = LdrToLinear(px[2]), o[3] = px[3] * 1.f / 255.f;
            }
            std::swap(prev, cur);
        }
        src += ph * (stride + 1);
    }
    return true;
}

// ---- OpenEXR, single-part scan-line files ------------------------------------------------------------------------------
float HalfToFloat(uint16_t hbits) {
    const uint32_t s = (hbits >> 15) & 1u, e = (hbits >> 10) & 0x1fu, m = hbits & 0x3ffu;
    uint32_t out;
    if (e == 0) {
        if (m == 0) out = s << 31;
        else { // subnormal
            int ee = -1;
            uint32_t mm = m;
            do {
                ++ee;
                mm <<= 1;
            } while (!(mm & 0x400u));
            out = s << 31 | uint32_t(127 - 15 - ee) << 23 | (mm & 0x3ffu) << 13;
        }
    } else if (e == 31) {
        out = s << 31 | 0x7f800000u | m << 13;
    } else {
        out = s << 31 | (e + 127 - 15) << 23 | m << 13;
    }
    float f;
    std::memcpy(&f, &out, 4);
    return f;
}
// the reordering + predictor OpenEXR applies around zlib / RLE blocks
void ExrUnpredict(std::vector<uint8_t> &t, std::vector<uint8_t> &out) {
    for (size_t i = 1; i < t.size(); ++i) t[i] = static_cast<uint8_t>(t[i - 1] + t[i] - 128);
    out.resize(t.size());
    const size_t half = (t.size() + 1) / 2;
    for (size_t i = 0; i < t.size(); ++i) out[i] = (i & 1) ? t[half + i / 2] : t[i / 2];
}
void ExrPredict(const uint8_t *in, size_t n, std::vector<uint8_t> &t) {
    t.resize(n);
    const size_t half = (n + 1) / 2;
    for (size_t i = 0; i < n; ++i) t[(i & 1) ? half + i / 2 : i / 2] = in[i];
    uint8_t prev = t.empty() ? 0 : t[0];
    for (size_t i = 1; i < n; ++i) {
        const uint8_t cur = t[i];
        t[i] = static_cast<uint8_t>(cur - prev + 128);
        prev = cur;
    }
}
bool ExrRleDecode(const uint8_t *in, size_t n, std::vector<uint8_t> &out, size_t expect) {
    out.clear();
    size_t i = 0;
    while (i < n) {
        const int8_t c = static_cast<int8_t>(in[i++]);
        if (c < 0) {
            const size_t cnt = static_cast<size_t>(-c);
            if (i + cnt > n) return false;
            out.insert(out.end(), in + i, in + i + cnt);
            i += cnt;
        } else {
            if (i >= n) return false;
            out.insert(out.end(), static_cast<size_t>(c) + 1, in[i++]);
        }
    }
    return out.size() == expect;
}
struct ExrChannel {
    std::string name;
    int type = 2; // 0 UINT, 1 HALF, 2 FLOAT
    size_t bytes() const { return type == 1 ? 2 : 4; }
};
bool LoadExr(const std::vector<uint8_t> &file, Image &img, std::string &why) {
    Reader r(file.data(), file.size());
    if (r.le<uint32_t>() != 20000630u) return false;
    const uint32_t version = r.le<uint32_t>();
    if ((version & 0xffu) != 2 || (version & 0x1800u)) { // deep data or multi-part
        why = "only single-part flat EXR files are read";
        return false;
    }
    const bool tiled = (version & 0x200u) != 0;
    uint32_t tile_w = 0, tile_h = 0;
    int tile_mode = -1;
    std::vector<ExrChannel> channels;
    int compression = -1, line_order = 0;
    int dw[4] = { 0, 0, -1, -1 };
    for (;;) {
        const std::string name = r.cstr();
        if (!r.ok) return false;
        if (name.empty()) break;
        const std::string type = r.cstr();
        const uint32_t size = r.le<uint32_t>();
        const uint8_t *body = r.take(size);
        if (!r.ok || !body) return false;
        if (name == "channels") {
            Reader c(body, size);
            for (;;) {
                ExrChannel ch;
                ch.name = c.cstr();
                if (ch.name.empty()) break;
                ch.type = c.le<int32_t>();
                c.take(4);
                const int xs = c.le<int32_t>(), ys = c.le<int32_t>();
                if (!c.ok || xs != 1 || ys != 1) {
                    why = "subsampled channels";
                    return false;
                }
                channels.push_back(ch);
            }
        } else if (name == "compression" && size >= 1) {
            compression = body[0];
        } else if (name == "dataWindow" && size >= 16) {
            std::memcpy(dw, body, 16);
        } else if (name == "lineOrder" && size >= 1) {
            line_order = body[0];
        } else if (name == "tiles" && size >= 9) {
            std::memcpy(&tile_w, body, 4), std::memcpy(&tile_h, body + 4, 4);
            tile_mode = body[8];
        }
    }
    if (channels.empty() || dw[2] < dw[0] || dw[3] < dw[1]) return false;
    int lines_per_block;
    switch (compression) {
        case 0: case 1: case 2: lines_per_block = 1; break;
        case 3: lines_per_block = 16; break;
        case 4: lines_per_block = 32; break;
        default: why = "compression " + std::to_string(compression) + " (only NONE, RLE, ZIPS, ZIP and PIZ are read)"; return false;
    }
    (void)line_order; // chunks carry their own y; the offset table is ignored and chunks are read in file order
    const size_t w = static_cast<size_t>(dw[2] - dw[0] + 1), h = static_cast<size_t>(dw[3] - dw[1] + 1);
    if (tiled && (tile_mode < 0 || (tile_mode & 0xf) != 0 || !tile_w || !tile_h)) {
        why = "only single-level (ONE_LEVEL) tiled files are read";
        return false;
    }
    const size_t tiles_x = tiled ? (w + tile_w - 1) / tile_w : 0, tiles_y = tiled ? (h + tile_h - 1) / tile_h : 0;
    const size_t n_blocks = tiled ? tiles_x * tiles_y : (h + lines_per_block - 1) / lines_per_block;
    if (n_blocks > file.size() / 8) return false;
    r.take(n_blocks * 8);
    size_t pixel_bytes = 0;
    for (auto &c : channels) pixel_bytes += c.bytes();
    const size_t line_bytes = pixel_bytes * w;
    // channel -> RGBA slot; a single channel (luminance) is replicated like tinyexr's LoadEXR does
    auto slot_of = [&](const std::string &n) { return n == "R" ? 0 : n == "G" ? 1 : n == "B" ? 2 : n == "A" ? 3 : -1; };
    // the offset table must fit, and zlib / RLE expand by at most ~1032 : 1: a data window the file's bytes cannot fill is corrupt
    // (and must not be able to ask for gigabytes)
    if (!r.ok || dw[2] < dw[0] || dw[3] < dw[1] || w > (1u << 20) || h > (1u << 20) || w * h > (size_t(1) << 28) || line_bytes * h > file.size() * 1040 + 64) return false;
    img.w = w, img.h = h;
    img.rgba.assign(w * h * 4, 0.f);
    for (size_t i = 0; i < w * h; ++i) img.rgba[i * 4 + 3] = 1.f;
    std::vector<uint8_t> tmp, block;
    // one chunk (a block of scan lines, or a tile): cw x lines pixels whose top-left corner is (x0, first) of the picture
    auto chunk = [&](const uint8_t *body, uint32_t size, size_t x0, size_t first, size_t cw, size_t lines) -> bool {
        const size_t expect = lines * pixel_bytes * cw;
        if (compression == 0 || size == expect) { // stored raw (also what writers do when compression does not pay)
            if (size != expect) return false;
            block.assign(body, body + size);
        } else if (compression == 1) {
            if (!ExrRleDecode(body, size, tmp, expect)) return false;
            ExrUnpredict(tmp, block);
        } else if (compression == 4) {
            std::vector<int> words;
            for (auto &c : channels) words.push_back(static_cast<int>(c.bytes() / 2));
            if (!piz::Decompress(body, size, cw, lines, words, block) || block.size() != expect) return false;
        } else {
            tmp.resize(expect);
            uLongf len = static_cast<uLongf>(expect);
            if (uncompress(tmp.data(), &len, body, size) != Z_OK || len != expect) return false;
            ExrUnpredict(tmp, block);
        }
        const uint8_t *p = block.data();
        for (size_t l = 0; l < lines; ++l)
            for (auto &c : channels) {
                const int slot = channels.size() == 1 ? 4 : slot_of(c.name);
                for (size_t x = 0; x < cw; ++x, p += c.bytes()) {
                    float v;
                    if (c.type == 1) {
                        uint16_t hv;
                        std::memcpy(&hv, p, 2);
                        v = HalfToFloat(hv);
                    } else if (c.type == 2) {
                        std::memcpy(&v, p, 4);
                    } else {
                        uint32_t u;
                        std::memcpy(&u, p, 4);
                        v = static_cast<float>(u);
                    }
                    float *o = &img.rgba[((first + l) * w + x0 + x) * 4];
                    if (slot == 4) o[0] = o[1] = o[2] = o[3] = v; // LoadEXR copies a lone channel into all four, alpha included
                    else if (slot >= 0) o[slot] = v;
                }
            }
        return true;
    };
    for (size_t b = 0; b < n_blocks; ++b) {
        if (tiled) { // tile coordinates, level (0, 0), size, data
            const int tx = r.le<int32_t>(), ty = r.le<int32_t>(), lx = r.le<int32_t>(), ly = r.le<int32_t>();
            const uint32_t size = r.le<uint32_t>();
            const uint8_t *body = r.take(size);
            if (!r.ok || !body || tx < 0 || ty < 0 || static_cast<size_t>(tx) >= tiles_x || static_cast<size_t>(ty) >= tiles_y || lx != 0 || ly != 0) return false;
            const size_t x0 = static_cast<size_t>(tx) * tile_w, y0 = static_cast<size_t>(ty) * tile_h;
            if (!chunk(body, size, x0, y0, std::min<size_t>(tile_w, w - x0), std::min<size_t>(tile_h, h - y0))) return false;
        } else {
            const int y0 = r.le<int32_t>();
            const uint32_t size = r.le<uint32_t>();
            const uint8_t *body = r.take(size);
            if (!r.ok || !body || y0 < dw[1] || y0 > dw[3]) return false;
            const size_t first = static_cast<size_t>(y0 - dw[1]);
            if (!chunk(body, size, 0, first, w, std::min<size_t>(lines_per_block, h - first))) return false;
        }
    }
    return true;
}
template<typename T>
void Put(std::vector<uint8_t> &o, T v) {
    const uint8_t *p = reinterpret_cast<const uint8_t *>(&v);
    o.insert(o.end(), p, p + sizeof(T));
}
void PutStr(std::vector<uint8_t> &o, const char *s) { o.insert(o.end(), s, s + std::strlen(s) + 1); }
void PutAttr(std::vector<uint8_t> &o, const char *name, const char *type, const std::vector<uint8_t> &body) {
    PutStr(o, name), PutStr(o, type);
    Put<uint32_t>(o, static_cast<uint32_t>(body.size()));
    o.insert(o.end(), body.begin(), body.end());
}
// texture.cpp:23-85: three FLOAT channels named B, G, R, picture flipped, tinyexr's default ZIP compression
bool SaveExr(const float *data, size_t w, size_t h, const std::string &path) {
    std::vector<uint8_t> o;
    Put<uint32_t>(o, 20000630u), Put<uint32_t>(o, 2u);
    std::vector<uint8_t> a;
    for (const char *n : { "B", "G", "R" }) {
        PutStr(a, n);
        Put<int32_t>(a, 2), Put<uint32_t>(a, 0u), Put<int32_t>(a, 1), Put<int32_t>(a, 1);
    }
    a.push_back(0);
    PutAttr(o, "channels", "chlist", a);
    PutAttr(o, "compression", "compression", { 3 });
    a.clear();
    Put<int32_t>(a, 0), Put<int32_t>(a, 0), Put<int32_t>(a, static_cast<int32_t>(w) - 1), Put<int32_t>(a, static_cast<int32_t>(h) - 1);
    PutAttr(o, "dataWindow", "box2i", a);
    PutAttr(o, "displayWindow", "box2i", a);
    PutAttr(o, "lineOrder", "lineOrder", { 0 });
    a.clear();
    Put<float>(a, 1.f);
    PutAttr(o, "pixelAspectRatio", "float", a);
    a.clear();
    Put<float>(a, 0.f), Put<float>(a, 0.f);
    PutAttr(o, "screenWindowCenter", "v2f", a);
    a.clear();
    Put<float>(a, 1.f);
    PutAttr(o, "screenWindowWidth", "float", a);
    o.push_back(0);
    const size_t n_blocks = (h + 15) / 16, table_at = o.size();
    o.resize(o.size() + n_blocks * 8);
    std::vector<uint8_t> raw, pred, comp;
    for (size_t b = 0; b < n_blocks; ++b) {
        const uint64_t offset = o.size();
        std::memcpy(&o[table_at + b * 8], &offset, 8);
        const size_t y0 = b * 16, lines = std::min<size_t>(16, h - y0);
        raw.clear();
        for (size_t l = 0; l < lines; ++l) {
            const float *row = data + (h - 1 - (y0 + l)) * w * 4; // flip
            for (int ch : { 2, 1, 0 })                               // B, G, R
                for (size_t x = 0; x < w; ++x) Put<float>(raw, row[x * 4 + ch]);
        }
        ExrPredict(raw.data(), raw.size(), pred);
        uLongf clen = compressBound(static_cast<uLong>(pred.size()));
        comp.resize(clen);
        if (compress(comp.data(), &clen, pred.data(), static_cast<uLong>(pred.size())) != Z_OK) return false;
        Put<int32_t>(o, static_cast<int32_t>(y0));
        if (clen < raw.size()) {
            Put<uint32_t>(o, static_cast<uint32_t>(clen));
            o.insert(o.end(), comp.begin(), comp.begin() + clen);
        } else {
            Put<uint32_t>(o, static_cast<uint32_t>(raw.size()));
            o.insert(o.end(), raw.begin(), raw.end());
        }
    }
    std::FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = std::fwrite(o.data(), 1, o.size(), f) == o.size();
    return std::fclose(f) == 0 && ok;
}
// 8-bit RGB PNG of the displayed picture: filter type 0 on every row, one zlib stream
bool SavePng(const float *data, size_t w, size_t h, const std::string &path, DisplayTransform display) {
    std::vector<uint8_t> raw(h * (1 + w * 3));
    for (size_t y = 0; y < h; ++y) {
        uint8_t *row = &raw[y * (1 + w * 3)];
        *row++ = 0;
        const float *src = data + (h - 1 - y) * w * 4; // flip: buffer row 0 is the bottom of the picture
        for (size_t x = 0; x < w; ++x) {
            float c[3];
            DisplayColor(src + x * 4, display, c);
            for (int k = 0; k < 3; ++k) { // UNORM render target: clamp, scale, round to nearest
                const float v = c[k] < 0.f || c[k] != c[k] ? 0.f : (c[k] > 1.f ? 1.f : c[k]);
                *row++ = static_cast<uint8_t>(v * 255.f + 0.5f);
            }
        }
    }
    uLongf clen = compressBound(static_cast<uLong>(raw.size()));
    std::vector<uint8_t> comp(clen);
    if (compress(comp.data(), &clen, raw.data(), static_cast<uLong>(raw.size())) != Z_OK) return false;
    std::vector<uint8_t> o = { 0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a };
    auto be32 = [&](uint32_t v) { o.push_back(v >> 24), o.push_back(v >> 16 & 255), o.push_back(v >> 8 & 255), o.push_back(v & 255); };
    auto chunk = [&](const char *type, const uint8_t *body, size_t n) {
        be32(static_cast<uint32_t>(n));
        const size_t at = o.size();
        o.insert(o.end(), type, type + 4);
        o.insert(o.end(), body, body + n);
        be32(static_cast<uint32_t>(crc32(0L, o.data() + at, static_cast<uInt>(n + 4))));
    };
    uint8_t ihdr[13] = { uint8_t(w >> 24), uint8_t(w >> 16), uint8_t(w >> 8), uint8_t(w), uint8_t(h >> 24), uint8_t(h >> 16), uint8_t(h >> 8), uint8_t(h), 8, 2, 0, 0, 0 };
    chunk("IHDR", ihdr, 13);
    chunk("IDAT", comp.data(), clen);
    chunk("IEND", nullptr, 0);
    std::FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = std::fwrite(o.data(), 1, o.size(), f) == o.size();
    return std::fclose(f) == 0 && ok;
}
}// namespace

void DisplayColor(const float rgb_in[3], DisplayTransform display, float rgb_out[3]) noexcept {
    for (int k = 0; k < 3; ++k) {
        float c = rgb_in[k];
        if (display.tone_mapping) { // ACESToneMapping(color, adapted_lum = 1), output.hlsl:30-40
            const float A = 2.51f, B = 0.03f, Cc = 2.43f, D = 0.59f, E = 0.14f;
            c = (c * (A * c + B)) / (c * (Cc * c + D) + E);
        }
        if (display.gamma_correct) c = std::pow(c, 1.f / 2.2f); // GammaCorrection(color, 2.2), :42-48
        rgb_out[k] = c;
    }
}

bool LoadImage(std::string_view path, Image &out) noexcept {
    try {
        std::vector<uint8_t> file;
        if (!ReadFile(path, file)) {
            Log::Warn("fail to load image [%s]: cannot read the file", std::string(path).c_str());
            return false;
        }
        std::string why;
        bool ok = false;
        if (file.size() >= 4 && file[0] == 0x76 && file[1] == 0x2f && file[2] == 0x31 && file[3] == 0x01) ok = LoadExr(file, out, why);
        else if (file.size() >= 8 && file[0] == 0x89 && file[1] == 'P') ok = LoadPng(file, out);
        else if (file.size() >= 2 && file[0] == '#' && file[1] == '?') ok = LoadHdr(file, out);
        else if (file.size() >= 2 && file[0] == 'P' && (file[1] == 'F' || file[1] == 'f')) ok = LoadPfm(file, out);
        else {
            ldr::Pixels8 px;
            std::string ext = std::filesystem::path(std::string(path)).extension().string();
            std::transform(ext.begin(), ext.end(), ext.begin(), [](unsigned char ch) { return static_cast<char>(std::tolower(ch)); });
            if (file.size() >= 3 && file[0] == 0xff && file[1] == 0xd8 && file[2] == 0xff) ok = ldr::LoadJpeg(file.data(), file.size(), px, why);
            else if (file.size() >= 2 && file[0] == 'B' && file[1] == 'M') ok = ldr::LoadBmp(file.data(), file.size(), px, why);
            else if (file.size() >= 2 && file[0] == 'P' && (file[1] == '5' || file[1] == '6')) ok = ldr::LoadPnm(file.data(), file.size(), px, why);
            else if (ext == ".tga" && ldr::LooksLikeTga(file.data(), file.size())) ok = ldr::LoadTga(file.data(), file.size(), px, why); // no magic number
            else why = "unsupported format (hdr, exr, png, jpeg, bmp, tga, pgm / ppm and pfm are read)";
            if (ok) FromPixels8(px, out);
        }
        if (!ok || !out.Valid()) {
            Log::Warn("fail to load image [%s]%s%s", std::string(path).c_str(), why.empty() ? "" : ": ", why.c_str());
            out = Image{};
            return false;
        }
        Log::Info("load image [%s] (%zux%zu)", std::string(path).c_str(), out.w, out.h);
        return true;
    } catch (...) {
        out = Image{};
        return false;
    }
}

bool SaveImage(const float *data, size_t w, size_t h, std::string_view path, EImageFileFormat format, DisplayTransform display) noexcept {
    try {
        if (!data || !w || !h) return false;
        const std::string p(path);
        bool ok = false;
        switch (format) {
            case EImageFileFormat::HDR: ok = SaveHdr(data, w, h, p); break;
            case EImageFileFormat::EXR: ok = SaveExr(data, w, h, p); break;
            case EImageFileFormat::PFM: ok = SavePfm(data, w, h, p); break;
            case EImageFileFormat::PNG: ok = SavePng(data, w, h, p, display); break;
        }
        if (ok) Log::Info("image was saved successfully in [%s].", p.c_str());
        else Log::Warn("image saving failed.");
        return ok;
    } catch (...) {
        return false;
    }
}
}// namespace Pupil::util
