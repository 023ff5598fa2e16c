#include "image_ldr.h"

#include <algorithm>
#include <array>
#include <cstring>
#include <memory>

namespace Pupil::util::ldr {
namespace {
// =====================================================================================================================
// JPEG (ITU T.81): Huffman-coded DCT processes — baseline (SOF0), extended sequential (SOF1) and progressive (SOF2)
// =====================================================================================================================
constexpr uint8_t kZigZag[64] = { 0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                  41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                  30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63 };

struct HuffTable {
    bool defined = false;
    uint8_t symbols[256] = {};
    int first_code[17] = {}, last_code[17] = {}, first_index[17] = {}; // per code length 1..16; last_code = -1: none
    uint16_t fast[512] = {};                                           // 9-bit prefix -> (length << 8) | symbol, 0 = longer code
    bool Build(const uint8_t counts[16], const uint8_t *syms, int n_syms) {
        std::memset(fast, 0, sizeof fast);
        std::memcpy(symbols, syms, static_cast<size_t>(n_syms));
        int code = 0, index = 0;
        for (int len = 1; len <= 16; ++len) {
            const int n = counts[len - 1];
            first_index[len] = index, first_code[len] = code;
            last_code[len] = n ? code + n - 1 : -1;
            if (code + n > (1 << len)) return false; // over-subscribed
            if (len <= 9) {
                for (int i = 0; i < n; ++i) {
                    const int prefix = (code + i) << (9 - len);
                    for (int f = 0; f < (1 << (9 - len)); ++f) fast[prefix + f] = static_cast<uint16_t>((len << 8) | symbols[index + i]);
                }
            }
            code = (code + n) << 1, index += n;
        }
        return defined = true;
    }
};

// entropy-coded segment reader: MSB first, FF00 unstuffed, stops feeding (zeros) at the first marker
struct BitReader {
    const uint8_t *p, *end;
    uint32_t acc = 0; // valid bits at the top
    int count = 0;
    bool at_marker = false;
    void Fill() {
        while (count <= 24) {
            uint32_t b = 0;
            if (!at_marker && p < end) {
                b = *p;
                if (b == 0xff) {
                    if (p + 1 < end && p[1] == 0x00) p += 2;
                    else at_marker = true, b = 0; // p stays on the marker
                } else ++p;
            } else at_marker = true;
            acc |= b << (24 - count);
            count += 8;
        }
    }
    int Get(int n) { // n <= 16
        if (n == 0) return 0;
        if (count < n) Fill();
        const int v = static_cast<int>(acc >> (32 - n));
        acc <<= n, count -= n;
        return v;
    }
    int Bit() { return Get(1); }
    int Extend(int n) { // T.81 F.2.2.1 RECEIVE + EXTEND
        if (n == 0) return 0;
        const int v = Get(n);
        return v < (1 << (n - 1)) ? v - (1 << n) + 1 : v;
    }
    int Symbol(const HuffTable &h) {
        if (count < 16) Fill();
        const uint16_t f = h.fast[acc >> 23];
        if (f) {
            acc <<= (f >> 8), count -= (f >> 8);
            return f & 0xff;
        }
        for (int len = 10; len <= 16; ++len) {
            const int code = static_cast<int>(acc >> (32 - len));
            if (h.last_code[len] >= 0 && code <= h.last_code[len] && code >= h.first_code[len]) {
                acc <<= len, count -= len;
                return h.symbols[h.first_index[len] + code - h.first_code[len]];
            }
        }
        return -1;
    }
    // restart boundary: drop the partial byte, step over RSTn
    bool Restart() {
        acc = 0, count = 0, at_marker = false;
        while (p < end && *p != 0xff) ++p; // tolerate garbage before the marker
        while (p + 1 < end && p[0] == 0xff && p[1] == 0xff) ++p;
        if (p + 1 < end && p[0] == 0xff && p[1] >= 0xd0 && p[1] <= 0xd7) {
            p += 2;
            return true;
        }
        return false;
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0;
    int dc_table = 0, ac_table = 0;
    int blocks_w = 0, blocks_h = 0; // allocated: whole MCUs
    int px_w = 0, px_h = 0;         // samples the component really has: ceil(W * h / hmax), ceil(H * v / vmax)
    int dc_pred = 0;
    bool quant_latched = false;
    uint16_t quant[64] = {}; // zig-zag order, latched at the component's first scan
    std::vector<int16_t> coef; // blocks_w * blocks_h * 64, natural order, not dequantised
    std::vector<uint8_t> plane; // blocks_w * 8 x blocks_h * 8
};

// Fixed-point inverse DCT after Loeffler, Ligtenberg and Moschytz (the "slow integer" scheme of the IJG code), with the
// precision rules stb_image states for its decoder: constants rounded to 12 fractional bits as int(x * 4096 + 0.5), the
// column pass keeps 2 extra bits (>> 10 after + 512), the row pass removes 17 with the rounding term and the + 128 level
// shift folded into one bias.  Following those rules makes the texels equal to the reference's.
constexpr int Fx(float x) { return static_cast<int>(x * 4096 + 0.5); }
// (64-bit intermediates: corrupt files can carry coefficients that overflow 32 bits; valid ones give the same values either way)
using I64 = long long;
struct Lm8 {
    I64 even[4], odd[4]; // out[k] = even[k] + odd[3 - k], out[7 - k] = even[k] - odd[3 - k]
};
inline Lm8 LmButterfly(I64 s0, I64 s1, I64 s2, I64 s3, I64 s4, I64 s5, I64 s6, I64 s7, I64 bias) {
    Lm8 r;
    const I64 z = (s2 + s6) * Fx(0.5411961f);
    const I64 a2 = z + s6 * Fx(-1.847759065f), a3 = z + s2 * Fx(0.765366865f);
    const I64 a0 = (s0 + s4) * 4096 + bias, a1 = (s0 - s4) * 4096 + bias;
    r.even[0] = a0 + a3, r.even[3] = a0 - a3, r.even[1] = a1 + a2, r.even[2] = a1 - a2;
    const I64 q3 = s7 + s3, q4 = s5 + s1, q1 = s7 + s1, q2 = s5 + s3;
    const I64 q5 = (q3 + q4) * Fx(1.175875602f);
    const I64 m1 = q5 + q1 * Fx(-0.899976223f), m2 = q5 + q2 * Fx(-2.562915447f);
    const I64 m3 = q3 * Fx(-1.961570560f), m4 = q4 * Fx(-0.390180644f);
    r.odd[0] = s7 * Fx(0.298631336f) + m1 + m3;
    r.odd[1] = s5 * Fx(2.053119869f) + m2 + m4;
    r.odd[2] = s3 * Fx(3.072711026f) + m2 + m3;
    r.odd[3] = s1 * Fx(1.501321110f) + m1 + m4;
    return r;
}
inline uint8_t Clamp8(I64 x) { return static_cast<uint8_t>(x < 0 ? 0 : (x > 255 ? 255 : x)); }
void InverseDct(const int16_t *c, uint8_t *out, int stride) {
    I64 tmp[64];
    for (int x = 0; x < 8; ++x) {
        const Lm8 r = LmButterfly(c[x], c[8 + x], c[16 + x], c[24 + x], c[32 + x], c[40 + x], c[48 + x], c[56 + x], 512);
        for (int k = 0; k < 4; ++k) tmp[k * 8 + x] = (r.even[k] + r.odd[3 - k]) >> 10, tmp[(7 - k) * 8 + x] = (r.even[k] - r.odd[3 - k]) >> 10;
    }
    for (int y = 0; y < 8; ++y) {
        const I64 *t = tmp + y * 8;
        const Lm8 r = LmButterfly(t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], 65536 + (128 << 17));
        uint8_t *o = out + y * stride;
        for (int k = 0; k < 4; ++k) o[k] = Clamp8((r.even[k] + r.odd[3 - k]) >> 17), o[7 - k] = Clamp8((r.even[k] - r.odd[3 - k]) >> 17);
    }
}

struct JpegDecoder {
    const uint8_t *file;
    size_t size;
    std::string &why;
    int width = 0, height = 0, n_comp = 0, h_max = 1, v_max = 1, mcus_x = 0, mcus_y = 0;
    bool progressive = false, have_frame = false, jfif = false;
    int adobe_transform = -1, restart_interval = 0;
    uint16_t quant[4][64] = {};
    bool quant_defined[4] = {};
    HuffTable dc[4], ac[4];
    Component comp[4];
    int eob_run = 0;

    JpegDecoder(const uint8_t *f, size_t n, std::string &w) : file(f), size(n), why(w) {}
    bool Fail(const char *msg) {
        why = std::string("jpeg: ") + msg;
        return false;
    }
    static int Be16(const uint8_t *p) { return p[0] << 8 | p[1]; }

    bool ParseFrame(const uint8_t *p, int len, int marker) {
        if (have_frame) return Fail("more than one frame");
        if (len < 6) return Fail("short SOF");
        if (p[0] != 8) return Fail("only 8-bit samples are read");
        height = Be16(p + 1), width = Be16(p + 3), n_comp = p[5];
        if (!width || !height) return Fail("zero size (DNL-defined heights are not read)");
        if (n_comp != 1 && n_comp != 3 && n_comp != 4) return Fail("only 1-, 3- and 4-component pictures are read");
        if (len < 6 + 3 * n_comp) return Fail("short SOF");
        if (static_cast<uint64_t>(width) * height > (1ull << 28)) return Fail("picture too large");
        progressive = marker == 0xc2;
        for (int i = 0; i < n_comp; ++i) {
            Component &c = comp[i];
            c.id = p[6 + 3 * i], c.h = p[7 + 3 * i] >> 4, c.v = p[7 + 3 * i] & 15, c.tq = p[8 + 3 * i];
            if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) return Fail("bad component parameters");
            h_max = std::max(h_max, c.h), v_max = std::max(v_max, c.v);
        }
        for (int i = 0; i < n_comp; ++i)
            if (h_max % comp[i].h || v_max % comp[i].v) return Fail("fractional sampling ratios are not read");
        mcus_x = (width + 8 * h_max - 1) / (8 * h_max), mcus_y = (height + 8 * v_max - 1) / (8 * v_max);
        // every coded block costs at least one bit, so a header that promises more blocks than the file has bits is corrupt
        // (and must not be able to ask for gigabytes of coefficient storage)
        uint64_t blocks_total = 0;
        for (int i = 0; i < n_comp; ++i) blocks_total += static_cast<uint64_t>(mcus_x) * comp[i].h * mcus_y * comp[i].v;
        if (blocks_total > static_cast<uint64_t>(size) * 8) return Fail("frame larger than its data");
        for (int i = 0; i < n_comp; ++i) {
            Component &c = comp[i];
            c.blocks_w = mcus_x * c.h, c.blocks_h = mcus_y * c.v;
            c.px_w = (width * c.h + h_max - 1) / h_max, c.px_h = (height * c.v + v_max - 1) / v_max;
            c.coef.assign(static_cast<size_t>(c.blocks_w) * c.blocks_h * 64, 0);
        }
        return have_frame = true;
    }
    bool ParseQuant(const uint8_t *p, int len) {
        while (len > 0) {
            const int pq = p[0] >> 4, tq = p[0] & 15, bytes = pq ? 128 : 64;
            if (pq > 1 || tq > 3 || len < 1 + bytes) return Fail("bad DQT");
            for (int k = 0; k < 64; ++k) quant[tq][k] = static_cast<uint16_t>(pq ? Be16(p + 1 + 2 * k) : p[1 + k]);
            quant_defined[tq] = true;
            p += 1 + bytes, len -= 1 + bytes;
        }
        return true;
    }
    bool ParseHuffman(const uint8_t *p, int len) {
        while (len > 0) {
            if (len < 17) return Fail("bad DHT");
            const int tc = p[0] >> 4, th = p[0] & 15;
            int total = 0;
            for (int i = 0; i < 16; ++i) total += p[1 + i];
            if (tc > 1 || th > 3 || total > 256 || len < 17 + total) return Fail("bad DHT");
            if (!(tc ? ac[th] : dc[th]).Build(p + 1, p + 17, total)) return Fail("over-subscribed Huffman table");
            p += 17 + total, len -= 17 + total;
        }
        return true;
    }

    // ---- one block of one scan ----
    bool BlockSequential(BitReader &br, Component &c, int16_t *blk) {
        const int t = br.Symbol(dc[c.dc_table]);
        if (t < 0 || t > 15) return Fail("bad DC code");
        c.dc_pred += br.Extend(t);
        blk[0] = static_cast<int16_t>(c.dc_pred);
        const HuffTable &h = ac[c.ac_table];
        for (int k = 1; k < 64;) {
            const int rs = br.Symbol(h);
            if (rs < 0) return Fail("bad AC code");
            const int run = rs >> 4, size = rs & 15;
            if (size == 0) {
                if (run != 15) break; // end of block
                k += 16;
                continue;
            }
            k += run;
            if (k > 63) return Fail("coefficient run past the block");
            blk[kZigZag[k++]] = static_cast<int16_t>(br.Extend(size));
        }
        return true;
    }
    bool BlockDcProgressive(BitReader &br, Component &c, int16_t *blk, int ah, int al) {
        if (ah == 0) {
            const int t = br.Symbol(dc[c.dc_table]);
            if (t < 0 || t > 15) return Fail("bad DC code");
            c.dc_pred += br.Extend(t);
            blk[0] = static_cast<int16_t>(c.dc_pred * (1 << al));
        } else if (br.Bit()) {
            blk[0] = static_cast<int16_t>(blk[0] + (1 << al));
        }
        return true;
    }
    bool BlockAcFirst(BitReader &br, Component &c, int16_t *blk, int ss, int se, int al) {
        if (eob_run > 0) {
            --eob_run;
            return true;
        }
        const HuffTable &h = ac[c.ac_table];
        for (int k = ss; k <= se;) {
            const int rs = br.Symbol(h);
            if (rs < 0) return Fail("bad AC code");
            const int run = rs >> 4, size = rs & 15;
            if (size == 0) {
                if (run < 15) { // EOBn: this band ends here in this and the next eob_run blocks
                    eob_run = (1 << run) - 1;
                    if (run) eob_run += br.Get(run);
                    break;
                }
                k += 16;
            } else {
                k += run;
                if (k > se) return Fail("coefficient run past the band");
                blk[kZigZag[k++]] = static_cast<int16_t>(br.Extend(size) * (1 << al));
            }
        }
        return true;
    }
    // T.81 G.1.2.3: successive-approximation refinement of an AC band.  Coefficients that are already non-zero receive one
    // correction bit each as the scan passes over them; newly non-zero ones arrive as (run of still-zero coefficients, sign).
    bool BlockAcRefine(BitReader &br, Component &c, int16_t *blk, int ss, int se, int al) {
        const int plus = 1 << al, minus = -(1 << al);
        auto correct = [&](int16_t &v) {
            if (br.Bit() && (v & plus) == 0) v = static_cast<int16_t>(v + (v >= 0 ? plus : minus));
        };
        int k = ss;
        if (eob_run == 0) {
            const HuffTable &h = ac[c.ac_table];
            while (k <= se) {
                const int rs = br.Symbol(h);
                if (rs < 0) return Fail("bad AC code");
                int run = rs >> 4;
                const int size = rs & 15;
                int fresh = 0;
                if (size == 0) {
                    if (run < 15) {
                        eob_run = 1 << run;
                        if (run) eob_run += br.Get(run);
                        break;
                    } // run == 15: sixteen still-zero coefficients
                } else {
                    if (size != 1) return Fail("bad refinement code");
                    fresh = br.Bit() ? plus : minus;
                }
                while (k <= se) {
                    int16_t &v = blk[kZigZag[k++]];
                    if (v != 0) correct(v);
                    else if (run-- == 0) {
                        v = static_cast<int16_t>(fresh);
                        break;
                    }
                }
            }
        }
        if (eob_run > 0) { // the rest of the band only carries correction bits
            for (; k <= se; ++k) {
                int16_t &v = blk[kZigZag[k]];
                if (v != 0) correct(v);
            }
            --eob_run;
        }
        return true;
    }

    bool DecodeScan(const uint8_t *header, int len, const uint8_t *&cursor) {
        if (!have_frame) return Fail("scan before frame");
        const int ns = len >= 1 ? header[0] : 0;
        if (ns < 1 || ns > n_comp || len < 4 + 2 * ns) return Fail("bad SOS");
        Component *sc[4];
        for (int i = 0; i < ns; ++i) {
            sc[i] = nullptr;
            for (int j = 0; j < n_comp; ++j)
                if (comp[j].id == header[1 + 2 * i]) sc[i] = &comp[j];
            if (!sc[i]) return Fail("scan names an unknown component");
            sc[i]->dc_table = header[2 + 2 * i] >> 4, sc[i]->ac_table = header[2 + 2 * i] & 15;
            if (sc[i]->dc_table > 3 || sc[i]->ac_table > 3) return Fail("bad table selector");
            if (!sc[i]->quant_latched) {
                if (!quant_defined[sc[i]->tq]) return Fail("missing quantisation table");
                std::memcpy(sc[i]->quant, quant[sc[i]->tq], sizeof sc[i]->quant);
                sc[i]->quant_latched = true;
            }
        }
        const int ss = header[1 + 2 * ns], se = header[2 + 2 * ns], ah = header[3 + 2 * ns] >> 4, al = header[3 + 2 * ns] & 15;
        if (progressive) {
            if (ss > 63 || se > 63 || ss > se || ah > 13 || al > 13 || (ss == 0 && se != 0) || (ss > 0 && ns != 1)) return Fail("bad progressive scan parameters");
        } else if (ss != 0 || se != 63 || ah != 0 || al != 0) {
            return Fail("bad sequential scan parameters");
        }
        for (int i = 0; i < ns; ++i) {
            const bool need_dc = !progressive || ss == 0, need_ac = !progressive || ss > 0;
            if (need_dc && !(progressive && ah) && !dc[sc[i]->dc_table].defined) return Fail("missing DC Huffman table");
            if (need_ac && !ac[sc[i]->ac_table].defined) return Fail("missing AC Huffman table");
            sc[i]->dc_pred = 0;
        }
        eob_run = 0;
        BitReader br{ cursor, file + size };
        auto block = [&](Component &c, int bx, int by) -> bool {
            int16_t *blk = c.coef.data() + (static_cast<size_t>(by) * c.blocks_w + bx) * 64;
            if (!progressive) return BlockSequential(br, c, blk);
            if (ss == 0) return BlockDcProgressive(br, c, blk, ah, al);
            return ah == 0 ? BlockAcFirst(br, c, blk, ss, se, al) : BlockAcRefine(br, c, blk, ss, se, al);
        };
        // a one-component scan walks that component's own blocks (ceil(px / 8)); an interleaved scan walks whole MCUs
        const int units_x = ns == 1 ? (sc[0]->px_w + 7) / 8 : mcus_x, units_y = ns == 1 ? (sc[0]->px_h + 7) / 8 : mcus_y;
        int until_restart = restart_interval;
        for (int uy = 0; uy < units_y; ++uy)
            for (int ux = 0; ux < units_x; ++ux) {
                if (ns == 1) {
                    if (!block(*sc[0], ux, uy)) return false;
                } else {
                    for (int i = 0; i < ns; ++i)
                        for (int by = 0; by < sc[i]->v; ++by)
                            for (int bx = 0; bx < sc[i]->h; ++bx)
                                if (!block(*sc[i], ux * sc[i]->h + bx, uy * sc[i]->v + by)) return false;
                }
                if (restart_interval && --until_restart == 0 && !(uy == units_y - 1 && ux == units_x - 1)) {
                    if (!br.Restart()) return Fail("missing restart marker");
                    for (int i = 0; i < ns; ++i) sc[i]->dc_pred = 0;
                    eob_run = 0, until_restart = restart_interval;
                }
            }
        cursor = br.p;
        return true;
    }

    void Reconstruct() {
        for (int i = 0; i < n_comp; ++i) {
            Component &c = comp[i];
            const int stride = c.blocks_w * 8;
            c.plane.assign(static_cast<size_t>(stride) * c.blocks_h * 8, 0);
            int16_t deq[64];
            for (int by = 0; by < c.blocks_h; ++by)
                for (int bx = 0; bx < c.blocks_w; ++bx) {
                    const int16_t *blk = c.coef.data() + (static_cast<size_t>(by) * c.blocks_w + bx) * 64;
                    for (int k = 0; k < 64; ++k) deq[kZigZag[k]] = static_cast<int16_t>(blk[kZigZag[k]] * c.quant[k]);
                    InverseDct(deq, c.plane.data() + static_cast<size_t>(by) * 8 * stride + bx * 8, stride);
                }
        }
    }

    // ---- chroma upsampling: triangle filter for the 2x cases (3/4 near + 1/4 far), replication otherwise ----
    static void RowH2(uint8_t *out, const uint8_t *in, int w) {
        if (w == 1) {
            out[0] = out[1] = in[0];
            return;
        }
        out[0] = in[0];
        out[1] = static_cast<uint8_t>((in[0] * 3 + in[1] + 2) >> 2);
        int i = 1;
        for (; i < w - 1; ++i) {
            const int n = 3 * in[i] + 2;
            out[2 * i] = static_cast<uint8_t>((n + in[i - 1]) >> 2);
            out[2 * i + 1] = static_cast<uint8_t>((n + in[i + 1]) >> 2);
        }
        out[2 * i] = static_cast<uint8_t>((in[w - 2] * 3 + in[w - 1] + 2) >> 2);
        out[2 * i + 1] = in[w - 1];
    }
    static void RowV2(uint8_t *out, const uint8_t *near, const uint8_t *far, int w) {
        for (int i = 0; i < w; ++i) out[i] = static_cast<uint8_t>((3 * near[i] + far[i] + 2) >> 2);
    }
    static void RowHV2(uint8_t *out, const uint8_t *near, const uint8_t *far, int w) {
        if (w == 1) {
            out[0] = out[1] = static_cast<uint8_t>((3 * near[0] + far[0] + 2) >> 2);
            return;
        }
        int t1 = 3 * near[0] + far[0];
        out[0] = static_cast<uint8_t>((t1 + 2) >> 2);
        for (int i = 1; i < w; ++i) {
            const int t0 = t1;
            t1 = 3 * near[i] + far[i];
            out[2 * i - 1] = static_cast<uint8_t>((3 * t0 + t1 + 8) >> 4);
            out[2 * i] = static_cast<uint8_t>((3 * t1 + t0 + 8) >> 4);
        }
        out[2 * w - 1] = static_cast<uint8_t>((t1 + 2) >> 2);
    }

    bool Output(Pixels8 &out) {
        const bool rgb_ids = n_comp == 3 && comp[0].id == 'R' && comp[1].id == 'G' && comp[2].id == 'B';
        const bool direct_rgb = n_comp == 3 && (rgb_ids || (adobe_transform == 0 && !jfif));
        // four components are print colours (Adobe CMYK, or YCCK when the APP14 transform flag is 2) and come out as RGB
        const int n_out = n_comp == 4 ? 3 : n_comp;
        out.w = width, out.h = height, out.channels = n_out;
        out.data.assign(static_cast<size_t>(width) * height * n_out, 0);
        struct Walk {
            int hs, vs, step, row, line0, line1;
        } walk[4];
        std::vector<uint8_t> line[4];
        for (int i = 0; i < n_comp; ++i) {
            walk[i] = Walk{ h_max / comp[i].h, v_max / comp[i].v, (v_max / comp[i].v) >> 1, 0, 0, 0 };
            line[i].resize(static_cast<size_t>(width) + 2 * 8 * h_max);
        }
        for (int y = 0; y < height; ++y) {
            const uint8_t *src[4];
            for (int i = 0; i < n_comp; ++i) {
                Walk &k = walk[i];
                const Component &c = comp[i];
                const int stride = c.blocks_w * 8;
                const bool lower = k.step >= (k.vs >> 1);
                const uint8_t *near = c.plane.data() + static_cast<size_t>(lower ? k.line1 : k.line0) * stride;
                const uint8_t *far = c.plane.data() + static_cast<size_t>(lower ? k.line0 : k.line1) * stride;
                if (k.hs == 1 && k.vs == 1) src[i] = near;
                else {
                    uint8_t *o = line[i].data();
                    if (k.hs == 1 && k.vs == 2) RowV2(o, near, far, c.px_w);
                    else if (k.hs == 2 && k.vs == 1) RowH2(o, near, c.px_w);
                    else if (k.hs == 2 && k.vs == 2) RowHV2(o, near, far, c.px_w);
                    else
                        for (int x = 0; x < c.px_w; ++x)
                            for (int j = 0; j < k.hs; ++j) o[x * k.hs + j] = near[x];
                    src[i] = o;
                }
                if (++k.step >= k.vs) {
                    k.step = 0, k.line0 = k.line1;
                    if (++k.row < c.px_h) ++k.line1;
                }
            }
            uint8_t *o = out.data.data() + static_cast<size_t>(y) * width * n_out;
            auto mul255 = [](int a, int b) { // a * b / 255, rounded (the 8 x 8 bit product rule stb_image uses for CMYK)
                const int t = a * b + 128;
                return static_cast<uint8_t>((t + (t >> 8)) >> 8);
            };
            if (n_comp == 1) std::memcpy(o, src[0], static_cast<size_t>(width));
            else if (n_comp == 4 && adobe_transform == 0)
                for (int x = 0; x < width; ++x) {
                    const int k = src[3][x];
                    o[3 * x] = mul255(src[0][x], k), o[3 * x + 1] = mul255(src[1][x], k), o[3 * x + 2] = mul255(src[2][x], k);
                }
            else if (direct_rgb)
                for (int x = 0; x < width; ++x) o[3 * x] = src[0][x], o[3 * x + 1] = src[1][x], o[3 * x + 2] = src[2][x];
            else { // YCbCr -> RGB in 20-bit fixed point (coefficients rounded to 12 bits; the Cb term of green drops its low 16 bits)
                constexpr int kCrR = Fx(1.40200f) << 8, kCrG = -(Fx(0.71414f) << 8), kCbG = -(Fx(0.34414f) << 8), kCbB = Fx(1.77200f) << 8;
                for (int x = 0; x < width; ++x) {
                    const int yy = (src[0][x] << 20) + (1 << 19), cb = src[1][x] - 128, cr = src[2][x] - 128;
                    const int r = yy + cr * kCrR;
                    const int g = yy + cr * kCrG + static_cast<int>(static_cast<uint32_t>(cb * kCbG) & 0xffff0000u);
                    const int b = yy + cb * kCbB;
                    o[3 * x] = Clamp8(r >> 20), o[3 * x + 1] = Clamp8(g >> 20), o[3 * x + 2] = Clamp8(b >> 20);
                    if (n_comp == 4 && adobe_transform == 2) { // YCCK: the three colours are stored inverted, then scaled by K
                        const int k = src[3][x];
                        o[3 * x] = mul255(255 - o[3 * x], k), o[3 * x + 1] = mul255(255 - o[3 * x + 1], k), o[3 * x + 2] = mul255(255 - o[3 * x + 2], k);
                    }
                }
            }
        }
        return true;
    }

    bool Run(Pixels8 &out) {
        if (size < 4 || file[0] != 0xff || file[1] != 0xd8) return Fail("no SOI");
        const uint8_t *p = file + 2, *end = file + size;
        bool saw_scan = false;
        for (;;) {
            while (p < end && *p != 0xff) ++p;
            while (p < end && *p == 0xff) ++p;
            if (p >= end) break; // stb_image and libjpeg both accept a missing EOI once scans were read
            const int marker = *p++;
            if (marker == 0xd9) break;
            if (marker == 0x00 || marker == 0x01 || (marker >= 0xd0 && marker <= 0xd8)) continue; // stuffing, TEM, stray RSTn / SOI
            if (end - p < 2) return Fail("truncated segment");
            const int len = Be16(p) - 2;
            if (len < 0 || end - p - 2 < len) return Fail("truncated segment");
            const uint8_t *body = p + 2;
            p = body + len;
            switch (marker) {
                case 0xc0: case 0xc1: case 0xc2:
                    if (!ParseFrame(body, len, marker)) return false;
                    break;
                case 0xc3: case 0xc5: case 0xc6: case 0xc7: case 0xc9: case 0xca: case 0xcb: case 0xcd: case 0xce: case 0xcf: case 0xcc:
                    return Fail("lossless, hierarchical and arithmetic-coded JPEG are not read");
                case 0xc4:
                    if (!ParseHuffman(body, len)) return false;
                    break;
                case 0xdb:
                    if (!ParseQuant(body, len)) return false;
                    break;
                case 0xdd:
                    if (len < 2) return Fail("bad DRI");
                    restart_interval = Be16(body);
                    break;
                case 0xda:
                    if (!DecodeScan(body, len, p)) return false;
                    saw_scan = true;
                    break;
                case 0xe0:
                    if (len >= 5 && !std::memcmp(body, "JFIF", 5)) jfif = true;
                    break;
                case 0xee:
                    if (len >= 12 && !std::memcmp(body, "Adobe", 5)) adobe_transform = body[11];
                    break;
                default: break; // other APPn, COM, ...: skipped
            }
        }
        if (!have_frame || !saw_scan) return Fail("no image data");
        for (int i = 0; i < n_comp; ++i)
            if (!comp[i].quant_latched) return Fail("a component has no scan");
        Reconstruct();
        return Output(out);
    }
};

// =====================================================================================================================
// BMP
// =====================================================================================================================
uint32_t Le32(const uint8_t *p) { return p[0] | p[1] << 8 | p[2] << 16 | static_cast<uint32_t>(p[3]) << 24; }
uint32_t Le16(const uint8_t *p) { return p[0] | p[1] << 8; }
// a channel given by a bit mask, widened to 8 bits by bit replication (what stb_image does for 16 / 32-bit bitmaps)
struct MaskChannel {
    int shift = 0, bits = 0;
    explicit MaskChannel(uint32_t m) {
        if (!m) return;
        while (!((m >> shift) & 1u)) ++shift;
        while (shift + bits < 32 && ((m >> (shift + bits)) & 1u)) ++bits;
    }
    int Get(uint32_t v) const {
        if (!bits) return 0;
        uint32_t x = (v >> shift) & ((bits >= 32 ? 0u : (1u << bits)) - 1u);
        if (bits >= 8) return static_cast<int>(x >> (bits - 8));
        // replicate the pattern down to fill 8 bits
        int have = bits;
        uint32_t r = x << (8 - bits);
        while (have < 8) {
            r |= r >> have;
            have *= 2;
        }
        return static_cast<int>(r & 0xffu);
    }
};
}// namespace

bool LoadJpeg(const uint8_t *file, size_t n, Pixels8 &out, std::string &why) {
    auto dec = std::make_unique<JpegDecoder>(file, n, why);
    return dec->Run(out);
}

bool LoadBmp(const uint8_t *file, size_t n, Pixels8 &out, std::string &why) {
    auto fail = [&](const char *m) {
        why = std::string("bmp: ") + m;
        return false;
    };
    if (n < 26 || file[0] != 'B' || file[1] != 'M') return fail("no BM signature");
    const uint32_t data_offset = Le32(file + 10), hsz = Le32(file + 14);
    if (hsz != 12 && hsz != 40 && hsz != 56 && hsz != 108 && hsz != 124) return fail("unknown header size");
    if (n < 14 + static_cast<size_t>(hsz)) return fail("truncated header");
    int w, h, bpp;
    uint32_t compress = 0, mr = 0, mg = 0, mb = 0, ma = 0;
    if (hsz == 12) w = static_cast<int>(Le16(file + 18)), h = static_cast<int>(Le16(file + 20)), bpp = static_cast<int>(Le16(file + 24));
    else {
        w = static_cast<int32_t>(Le32(file + 18)), h = static_cast<int32_t>(Le32(file + 22)), bpp = static_cast<int>(Le16(file + 28));
        compress = Le32(file + 30);
    }
    const bool flip = h > 0; // positive height: rows are stored bottom-up
    h = h < 0 ? -h : h;
    if (w <= 0 || h <= 0 || static_cast<uint64_t>(w) * h > (1ull << 28)) return fail("bad size");
    if (compress == 1 || compress == 2) return fail("run-length encoded bitmaps are not read");
    if (compress != 0 && compress != 3) return fail("unknown compression");
    if (bpp == 16 || bpp == 32) {
        if (compress == 0) {
            if (bpp == 32) mr = 0xffu << 16, mg = 0xffu << 8, mb = 0xffu, ma = 0xffu << 24;
            else mr = 31u << 10, mg = 31u << 5, mb = 31u;
        } else {
            // BI_BITFIELDS: the masks follow a 40-byte header, or are part of the larger ones
            if (n < 14 + 40 + 12) return fail("truncated masks");
            mr = Le32(file + 54), mg = Le32(file + 58), mb = Le32(file + 62);
            if (hsz >= 56) ma = Le32(file + 66);
            if (!mr || !mg || !mb || (mr == mg && mg == mb)) return fail("bad masks");
        }
    } else if (compress != 0) return fail("bit fields need 16 or 32 bits per pixel");
    int palette_n = 0;
    const uint8_t *palette = file + 14 + hsz;
    const int palette_stride = hsz == 12 ? 3 : 4;
    if (bpp <= 8) {
        if (bpp != 1 && bpp != 4 && bpp != 8) return fail("unsupported bit depth");
        palette_n = hsz == 12 ? (static_cast<int>(data_offset) - 14 - 12) / 3 : static_cast<int>(Le32(file + 46));
        if (palette_n == 0 && hsz != 12) palette_n = (static_cast<int>(data_offset) - 14 - static_cast<int>(hsz)) / 4;
        if (palette_n <= 0 || palette_n > 256 || 14 + hsz + static_cast<size_t>(palette_n) * palette_stride > n) return fail("bad palette");
    } else if (bpp != 16 && bpp != 24 && bpp != 32) return fail("unsupported bit depth");
    const size_t stride = ((static_cast<size_t>(w) * bpp + 31) / 32) * 4;
    if (data_offset > n || n - data_offset < stride * (static_cast<size_t>(h) - 1) + (static_cast<size_t>(w) * bpp + 7) / 8) return fail("truncated pixel data");
    out.w = w, out.h = h, out.channels = ma ? 4 : 3;
    out.data.assign(static_cast<size_t>(w) * h * out.channels, 255);
    const MaskChannel cr(mr), cg(mg), cb(mb), ca(ma);
    uint32_t any_alpha = 0;
    for (int y = 0; y < h; ++y) {
        const uint8_t *row = file + data_offset + stride * static_cast<size_t>(y);
        uint8_t *o = out.data.data() + static_cast<size_t>(flip ? h - 1 - y : y) * w * out.channels;
        for (int x = 0; x < w; ++x, o += out.channels) {
            if (bpp <= 8) {
                const int per = 8 / bpp, idx = (row[x / per] >> (8 - bpp - (x % per) * bpp)) & ((1 << bpp) - 1);
                if (idx >= palette_n) return fail("palette index out of range");
                const uint8_t *c = palette + idx * palette_stride; // stored B, G, R
                o[0] = c[2], o[1] = c[1], o[2] = c[0];
            } else if (bpp == 24) {
                o[0] = row[3 * x + 2], o[1] = row[3 * x + 1], o[2] = row[3 * x];
            } else {
                const uint32_t v = bpp == 16 ? Le16(row + 2 * x) : Le32(row + 4 * x);
                o[0] = static_cast<uint8_t>(cr.Get(v)), o[1] = static_cast<uint8_t>(cg.Get(v)), o[2] = static_cast<uint8_t>(cb.Get(v));
                if (ma) o[3] = static_cast<uint8_t>(ca.Get(v)), any_alpha |= o[3];
            }
        }
    }
    // 32-bit BI_RGB files usually leave the fourth byte zero: a picture whose alpha is zero everywhere is opaque
    if (ma && any_alpha == 0)
        for (size_t i = 3; i < out.data.size(); i += 4) out.data[i] = 255;
    return true;
}

bool LoadPnm(const uint8_t *file, size_t n, Pixels8 &out, std::string &why) {
    auto fail = [&](const char *m) {
        why = std::string("pnm: ") + m;
        return false;
    };
    if (n < 3 || file[0] != 'P' || (file[1] != '5' && file[1] != '6')) return fail("only binary P5 / P6 files are read");
    size_t pos = 2;
    auto number = [&](int &v) -> bool { // white space and # comments, then decimal digits
        for (;;) {
            while (pos < n && (file[pos] == ' ' || file[pos] == '\t' || file[pos] == '\r' || file[pos] == '\n' || file[pos] == '\f' || file[pos] == '\v')) ++pos;
            if (pos < n && file[pos] == '#') {
                while (pos < n && file[pos] != '\n' && file[pos] != '\r') ++pos;
            } else break;
        }
        if (pos >= n || file[pos] < '0' || file[pos] > '9') return false;
        long long x = 0;
        while (pos < n && file[pos] >= '0' && file[pos] <= '9' && x < (1ll << 31)) x = x * 10 + (file[pos++] - '0');
        v = static_cast<int>(x < (1ll << 31) ? x : -1);
        return v >= 0;
    };
    int w = 0, h = 0, maxval = 0;
    if (!number(w) || !number(h) || !number(maxval)) return fail("bad header");
    if (w <= 0 || h <= 0 || maxval <= 0 || maxval > 65535 || static_cast<uint64_t>(w) * h > (1ull << 28)) return fail("bad header");
    if (pos >= n) return fail("truncated pixel data");
    ++pos; // the single white-space byte after maxval
    const int channels = file[1] == '6' ? 3 : 1, bytes = maxval > 255 ? 2 : 1;
    const size_t count = static_cast<size_t>(w) * h * channels;
    if ((n - pos) / bytes < count) return fail("truncated pixel data");
    out.w = w, out.h = h, out.channels = channels;
    out.data.resize(count);
    // 16-bit samples are stored big endian.  stb_image (the reference's reader) loads them as host words without swapping, so on
    // the little-endian hosts it runs on its 8-bit API hands back the SECOND byte of each sample; the same byte is taken here so
    // that such a texture yields the texels the reference would see
    for (size_t i = 0; i < count; ++i) out.data[i] = file[pos + i * bytes + (bytes - 1)];
    return true;
}

bool LooksLikeTga(const uint8_t *f, size_t n) {
    if (n < 18) return false;
    const int cmap = f[1], type = f[2], bpp = f[16];
    if (cmap > 1) return false;
    if (cmap == 1) {
        if (type != 1 && type != 9) return false;
        const int eb = f[7];
        if (eb != 8 && eb != 15 && eb != 16 && eb != 24 && eb != 32) return false;
    } else if (type != 2 && type != 3 && type != 10 && type != 11) return false;
    if (Le16(f + 12) < 1 || Le16(f + 14) < 1) return false;
    if (cmap == 1) return bpp == 8 || bpp == 16;
    return bpp == 8 || bpp == 15 || bpp == 16 || bpp == 24 || bpp == 32;
}

bool LoadTga(const uint8_t *file, size_t n, Pixels8 &out, std::string &why) {
    auto fail = [&](const char *m) {
        why = std::string("tga: ") + m;
        return false;
    };
    if (!LooksLikeTga(file, n)) return fail("not a readable TGA header");
    const int id_len = file[0], cmap = file[1];
    int type = file[2];
    const int cmap_first = static_cast<int>(Le16(file + 3)), cmap_len = static_cast<int>(Le16(file + 5)), cmap_bpp = file[7];
    const int w = static_cast<int>(Le16(file + 12)), h = static_cast<int>(Le16(file + 14)), bpp = file[16];
    const bool top_down = (file[17] >> 5) & 1;
    const bool rle = type >= 8;
    if (rle) type -= 8;
    const bool grey = type == 3;
    // bytes per stored element and channels out: 15/16-bit colour is 5-5-5 -> RGB; 16-bit grey is grey + alpha
    auto channels_of = [&](int bits, bool is_grey) -> int {
        switch (bits) {
            case 8: return 1;
            case 15: return 3;
            case 16: return is_grey ? 2 : 3;
            case 24: return 3;
            case 32: return 4;
            default: return 0;
        }
    };
    const int src_bits = cmap ? cmap_bpp : bpp, channels = channels_of(src_bits, grey && !cmap);
    if (!channels) return fail("unsupported bit depth");
    const bool rgb16 = (src_bits == 15 || src_bits == 16) && !(grey && !cmap);
    const size_t src_bytes = static_cast<size_t>((src_bits + 7) / 8), idx_bytes = static_cast<size_t>(bpp / 8);
    size_t pos = 18 + static_cast<size_t>(id_len);
    const uint8_t *palette = nullptr;
    if (cmap) {
        palette = file + pos;
        pos += static_cast<size_t>(cmap_len) * src_bytes;
        if (pos > n) return fail("truncated colour map");
    }
    auto put = [&](uint8_t *o, const uint8_t *s) { // one stored element -> channels bytes, colour stored B, G, R(, A)
        if (rgb16) {
            const uint32_t v = Le16(s);
            o[0] = static_cast<uint8_t>(((v >> 10) & 31u) * 255u / 31u), o[1] = static_cast<uint8_t>(((v >> 5) & 31u) * 255u / 31u),
            o[2] = static_cast<uint8_t>((v & 31u) * 255u / 31u);
        } else if (channels >= 3) {
            o[0] = s[2], o[1] = s[1], o[2] = s[0];
            if (channels == 4) o[3] = s[3];
        } else {
            o[0] = s[0];
            if (channels == 2) o[1] = s[1];
        }
    };
    const size_t elem = cmap ? idx_bytes : src_bytes, total = static_cast<size_t>(w) * h;
    // the header alone must not be able to ask for gigabytes: raw pixels need their bytes, a run-length packet (1 + elem bytes)
    // expands to at most 128 pixels
    if (pos > n || (!rle && (n - pos) / elem < total) || (rle && ((n - pos) / (1 + elem) + 1) * 128 < total)) return fail("truncated pixel data");
    out.w = w, out.h = h, out.channels = channels;
    out.data.assign(total * channels, 0);
    uint8_t px[4] = { 0, 0, 0, 0 };
    auto read_elem = [&]() -> bool {
        if (pos + elem > n) return false;
        if (cmap) {
            int idx = static_cast<int>(elem == 1 ? file[pos] : Le16(file + pos)) - cmap_first;
            if (idx < 0 || idx >= cmap_len) idx = 0;
            put(px, palette + static_cast<size_t>(idx) * src_bytes);
        } else put(px, file + pos);
        pos += elem;
        return true;
    };
    size_t i = 0;
    while (i < total) {
        size_t run = 1;
        bool repeat = false;
        if (rle) {
            if (pos >= n) return fail("truncated pixel data");
            const int c = file[pos++];
            run = static_cast<size_t>(c & 127) + 1, repeat = c & 128;
        }
        for (size_t k = 0; k < run && i < total; ++k, ++i) {
            if (k == 0 || !repeat)
                if (!read_elem()) return fail("truncated pixel data");
            const size_t y = i / w, x = i % w, oy = top_down ? y : static_cast<size_t>(h) - 1 - y;
            std::memcpy(out.data.data() + (oy * w + x) * channels, px, static_cast<size_t>(channels));
        }
    }
    return true;
}
}// namespace Pupil::util::ldr
