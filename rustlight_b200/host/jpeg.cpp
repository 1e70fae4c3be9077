// jpeg.cpp -- JPEG reader for Bitmap::read_ldr_image (structure.rs:649-668: image::open(..).to_rgb8() / 255), the format of the
// reference's own texture-light picture (`butterfly.jpg`, examples/cli.rs:424: progressive, 4:2:0) and of textures in PBRT scenes.
//
// Baseline / extended sequential (SOF0, SOF1) and progressive (SOF2) Huffman JPEG, 8 bit, 1 or 3 components, any sampling factors
// with h, v in {1, 2}, restart intervals.  The reference decodes through the `image` crate (jpeg-decoder), which is not vendored:
// what is restated here is the JPEG standard (ITU T.81) with the arithmetic of the IJG library's defaults -- the "slow" integer inverse
// DCT (jidctint), triangle-filter ("fancy") chroma upsampling and the 16-bit fixed-point YCbCr conversion -- so that the result can be
// pinned against an independent decoder (tests/test_jpeg.py compares with PIL = libjpeg-turbo, bit for bit).  jpeg-decoder's own integer
// IDCT differs from it by at most one level on some pixels: parity with the reference at that level is UNPINNED.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iterator>
#include <vector>

#include "rl_host.hpp"

namespace rlh {
namespace {

const int kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                         35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huff { // T.81 Annex C / F.2.2.3: canonical code tables
    bool set = false;
    uint8_t bits[17] = {0};
    uint8_t vals[256] = {0};
    int mincode[17], maxcode[18], valptr[17];
    void build() {
        int code = 0, k = 0;
        for (int l = 1; l <= 16; l++) {
            valptr[l] = k;
            mincode[l] = code;
            code += bits[l];
            k += bits[l];
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        set = true;
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0;
    int bw = 0, bh = 0;   // blocks per row / column as stored (padded to whole MCUs)
    int cw = 0, ch = 0;   // blocks per row / column that carry image data (non-interleaved scans)
    std::vector<int16_t> coef; // bw * bh * 64, natural (de-zigzagged) order
    int dc_pred = 0;
    int td = 0, ta = 0;
};

struct BitReader {
    const uint8_t *p, *end;
    uint32_t acc = 0;
    int cnt = 0;
    bool hit_marker = false;
    BitReader(const uint8_t *b, const uint8_t *e) : p(b), end(e) {}
    void fill() {
        while (cnt <= 24) {
            int byte = 0;
            if (!hit_marker && p < end) {
                byte = *p;
                if (byte == 0xFF) {
                    if (p + 1 < end && p[1] == 0x00) p += 2;
                    else { // a marker: feed zeros until the caller resynchronises
                        hit_marker = true;
                        byte = 0;
                    }
                } else p++;
            }
            acc |= (uint32_t)byte << (24 - cnt);
            cnt += 8;
        }
    }
    int bit() {
        if (cnt < 1) fill();
        const int b = (int)(acc >> 31);
        acc <<= 1, cnt--;
        return b;
    }
    int bits(int n) {
        if (n == 0) return 0;
        if (cnt < n) fill();
        const int v = (int)(acc >> (32 - n));
        acc <<= n, cnt -= n;
        return v;
    }
    void reset() { acc = 0, cnt = 0, hit_marker = false; }
};

int decode_huff(BitReader &br, const Huff &h) {
    int code = 0;
    for (int l = 1; l <= 16; l++) {
        code = (code << 1) | br.bit();
        if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) return h.vals[h.valptr[l] + code - h.mincode[l]];
    }
    throw Error("jpeg: bad Huffman code");
}
int extend(int v, int t) { return t == 0 ? 0 : (v < (1 << (t - 1)) ? v - (1 << t) + 1 : v); } // T.81 F.2.2.1

// IJG jidctint.c (the accurate integer inverse DCT, CONST_BITS = 13, PASS1_BITS = 2): dequantised coefficients -> samples 0..255
inline int descale(long x, int n) { return (int)((x + (1L << (n - 1))) >> n); }
void idct_islow(const int16_t *coef, const uint16_t *q, uint8_t *out, int stride) {
    const long F_0_298 = 2446, F_0_390 = 3196, F_0_541 = 4433, F_0_765 = 6270, F_0_899 = 7373, F_1_175 = 9633, F_1_501 = 12299, F_1_847 = 15137,
               F_1_961 = 16069, F_2_053 = 16819, F_2_562 = 20995, F_3_072 = 25172;
    const int CB = 13, P1 = 2;
    int ws[64];
    for (int c = 0; c < 8; c++) { // pass 1: columns
        const int16_t *in = coef + c;
        const uint16_t *qq = q + c;
        int *w = ws + c;
        if (in[8] == 0 && in[16] == 0 && in[24] == 0 && in[32] == 0 && in[40] == 0 && in[48] == 0 && in[56] == 0) {
            const int dc = (int)((long)in[0] * qq[0]) << P1;
            for (int r = 0; r < 8; r++) w[8 * r] = dc;
            continue;
        }
        long z2 = (long)in[16] * qq[16], z3 = (long)in[48] * qq[48];
        long z1 = (z2 + z3) * F_0_541;
        long tmp2 = z1 + z3 * (-F_1_847), tmp3 = z1 + z2 * F_0_765;
        z2 = (long)in[0] * qq[0], z3 = (long)in[32] * qq[32];
        long tmp0 = (z2 + z3) << CB, tmp1 = (z2 - z3) << CB;
        long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = (long)in[56] * qq[56], tmp1 = (long)in[40] * qq[40], tmp2 = (long)in[24] * qq[24], tmp3 = (long)in[8] * qq[8];
        z1 = tmp0 + tmp3, z2 = tmp1 + tmp2, z3 = tmp0 + tmp2;
        long z4 = tmp1 + tmp3, z5 = (z3 + z4) * F_1_175;
        tmp0 *= F_0_298, tmp1 *= F_2_053, tmp2 *= F_3_072, tmp3 *= F_1_501;
        z1 *= -F_0_899, z2 *= -F_2_562, z3 *= -F_1_961, z4 *= -F_0_390;
        z3 += z5, z4 += z5;
        tmp0 += z1 + z3, tmp1 += z2 + z4, tmp2 += z2 + z3, tmp3 += z1 + z4;
        w[0] = descale(tmp10 + tmp3, CB - P1), w[56] = descale(tmp10 - tmp3, CB - P1);
        w[8] = descale(tmp11 + tmp2, CB - P1), w[48] = descale(tmp11 - tmp2, CB - P1);
        w[16] = descale(tmp12 + tmp1, CB - P1), w[40] = descale(tmp12 - tmp1, CB - P1);
        w[24] = descale(tmp13 + tmp0, CB - P1), w[32] = descale(tmp13 - tmp0, CB - P1);
    }
    auto clamp = [](int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); };
    for (int r = 0; r < 8; r++) { // pass 2: rows, + 128 level shift
        const int *w = ws + 8 * r;
        uint8_t *o = out + (size_t)r * stride;
        long z2 = w[2], z3 = w[6];
        long z1 = (z2 + z3) * F_0_541;
        long tmp2 = z1 + z3 * (-F_1_847), tmp3 = z1 + z2 * F_0_765;
        long tmp0 = ((long)w[0] + w[4]) << CB, tmp1 = ((long)w[0] - w[4]) << CB;
        long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = w[7], tmp1 = w[5], tmp2 = w[3], tmp3 = w[1];
        z1 = tmp0 + tmp3, z2 = tmp1 + tmp2, z3 = tmp0 + tmp2;
        long z4 = tmp1 + tmp3, z5 = (z3 + z4) * F_1_175;
        tmp0 *= F_0_298, tmp1 *= F_2_053, tmp2 *= F_3_072, tmp3 *= F_1_501;
        z1 *= -F_0_899, z2 *= -F_2_562, z3 *= -F_1_961, z4 *= -F_0_390;
        z3 += z5, z4 += z5;
        tmp0 += z1 + z3, tmp1 += z2 + z4, tmp2 += z2 + z3, tmp3 += z1 + z4;
        const int S = CB + P1 + 3;
        o[0] = clamp(descale(tmp10 + tmp3, S) + 128), o[7] = clamp(descale(tmp10 - tmp3, S) + 128);
        o[1] = clamp(descale(tmp11 + tmp2, S) + 128), o[6] = clamp(descale(tmp11 - tmp2, S) + 128);
        o[2] = clamp(descale(tmp12 + tmp1, S) + 128), o[5] = clamp(descale(tmp12 - tmp1, S) + 128);
        o[3] = clamp(descale(tmp13 + tmp0, S) + 128), o[4] = clamp(descale(tmp13 - tmp0, S) + 128);
    }
}

struct Decoder {
    std::vector<uint8_t> file;
    uint16_t qt[4][64];
    bool qt_set[4] = {false, false, false, false};
    Huff dc[4], ac[4];
    std::vector<Component> comps;
    int width = 0, height = 0, hmax = 1, vmax = 1, mcux = 0, mcuy = 0;
    bool progressive = false, have_sof = false;
    int restart_interval = 0;
    int eobrun = 0;

    uint16_t be16(size_t at) const {
        if (at + 2 > file.size()) throw Error("jpeg: truncated file");
        return (uint16_t)((file[at] << 8) | file[at + 1]);
    }
    void parse_dqt(size_t at, size_t len) {
        size_t p = at;
        while (p < at + len) {
            const int pq = file[p] >> 4, tq = file[p] & 15;
            p++;
            if (tq > 3 || pq > 1) throw Error("jpeg: bad quantisation table");
            for (int i = 0; i < 64; i++) {
                qt[tq][kZigzag[i]] = pq ? be16(p) : file[p];
                p += pq ? 2 : 1;
            }
            qt_set[tq] = true;
        }
    }
    void parse_dht(size_t at, size_t len) {
        size_t p = at;
        while (p < at + len) {
            const int tc = file[p] >> 4, th = file[p] & 15;
            p++;
            if (tc > 1 || th > 3) throw Error("jpeg: bad Huffman table id");
            Huff &h = tc ? ac[th] : dc[th];
            int n = 0;
            for (int l = 1; l <= 16; l++) h.bits[l] = file[p + l - 1], n += h.bits[l];
            p += 16;
            if (n > 256 || p + n > at + len) throw Error("jpeg: bad Huffman table");
            std::memcpy(h.vals, &file[p], n);
            p += n;
            h.build();
        }
    }
    void parse_sof(size_t at, int marker) {
        if (have_sof) throw Error("jpeg: more than one frame");
        if (marker != 0xC0 && marker != 0xC1 && marker != 0xC2) throw Error("jpeg: only baseline, extended-sequential and progressive Huffman JPEG are read (no arithmetic / lossless / hierarchical coding)");
        progressive = marker == 0xC2;
        if (file[at] != 8) throw Error("jpeg: only 8-bit samples are read");
        height = be16(at + 1), width = be16(at + 3);
        const int nc = file[at + 5];
        if (width == 0 || height == 0 || (nc != 1 && nc != 3)) throw Error("jpeg: only grey and 3-component images are read");
        comps.resize(nc);
        for (int i = 0; i < nc; i++) {
            Component &c = comps[i];
            c.id = file[at + 6 + 3 * i], c.h = file[at + 7 + 3 * i] >> 4, c.v = file[at + 7 + 3 * i] & 15, c.tq = file[at + 8 + 3 * i];
            if (c.h < 1 || c.h > 2 || c.v < 1 || c.v > 2 || c.tq > 3) throw Error("jpeg: sampling factors other than 1 and 2 are not read");
            hmax = std::max(hmax, c.h), vmax = std::max(vmax, c.v);
        }
        if (nc == 1) comps[0].h = comps[0].v = hmax = vmax = 1; // a single component is never interleaved: its factors do not matter
        mcux = (width + 8 * hmax - 1) / (8 * hmax), mcuy = (height + 8 * vmax - 1) / (8 * vmax);
        for (Component &c : comps) {
            c.bw = mcux * c.h, c.bh = mcuy * c.v;
            const int sw = (width * c.h + hmax - 1) / hmax, sh = (height * c.v + vmax - 1) / vmax; // component size in samples
            c.cw = (sw + 7) / 8, c.ch = (sh + 7) / 8;
            c.coef.assign((size_t)c.bw * c.bh * 64, 0);
        }
        have_sof = true;
    }

    // ---- one block of one scan (T.81 F.2.2, G.1.2) ----
    void block_sequential(BitReader &br, Component &c, int16_t *b) {
        const int t = decode_huff(br, dc[c.td]);
        if (t > 11) throw Error("jpeg: bad DC size");
        c.dc_pred += extend(br.bits(t), t);
        b[0] = (int16_t)c.dc_pred;
        for (int k = 1; k < 64;) {
            const int rs = decode_huff(br, ac[c.ta]), r = rs >> 4, s = rs & 15;
            if (s == 0) {
                if (r != 15) break;
                k += 16;
                continue;
            }
            k += r;
            if (k > 63) throw Error("jpeg: AC index out of range");
            b[kZigzag[k]] = (int16_t)extend(br.bits(s), s);
            k++;
        }
    }
    void block_dc_first(BitReader &br, Component &c, int16_t *b, int al) {
        const int t = decode_huff(br, dc[c.td]);
        if (t > 11) throw Error("jpeg: bad DC size");
        c.dc_pred += extend(br.bits(t), t);
        b[0] = (int16_t)(c.dc_pred * (1 << al));
    }
    void block_dc_refine(BitReader &br, int16_t *b, int al) {
        if (br.bit()) b[0] = (int16_t)(b[0] | (1 << al));
    }
    void block_ac_first(BitReader &br, Component &c, int16_t *b, int ss, int se, int al) {
        if (eobrun > 0) {
            eobrun--;
            return;
        }
        for (int k = ss; k <= se;) {
            const int rs = decode_huff(br, ac[c.ta]), r = rs >> 4, s = rs & 15;
            if (s == 0) {
                if (r < 15) {
                    eobrun = (1 << r) - 1;
                    if (r) eobrun += br.bits(r);
                    break;
                }
                k += 16;
                continue;
            }
            k += r;
            if (k > se) throw Error("jpeg: AC index out of range");
            b[kZigzag[k]] = (int16_t)(extend(br.bits(s), s) * (1 << al));
            k++;
        }
    }
    void block_ac_refine(BitReader &br, Component &c, int16_t *b, int ss, int se, int al) { // T.81 G.1.2.3
        const int p1 = 1 << al, m1 = -(1 << al);
        int k = ss;
        if (eobrun == 0) {
            for (; k <= se;) {
                const int rs = decode_huff(br, ac[c.ta]);
                int r = rs >> 4;
                const int s = rs & 15;
                int val = 0;
                if (s == 0) {
                    if (r < 15) {
                        eobrun = (1 << r);
                        if (r) eobrun += br.bits(r);
                        break;
                    }
                } else {
                    if (s != 1) throw Error("jpeg: bad refinement scan");
                    val = br.bit() ? p1 : m1;
                }
                while (k <= se) { // skip r zero-history coefficients, refining the non-zero ones on the way
                    int16_t &z = b[kZigzag[k]];
                    if (z != 0) {
                        if (br.bit() && (z & p1) == 0) z = (int16_t)(z >= 0 ? z + p1 : z + m1);
                    } else {
                        if (r == 0) {
                            if (val) z = (int16_t)val;
                            k++;
                            break;
                        }
                        r--;
                    }
                    k++;
                }
            }
        }
        if (eobrun > 0) { // the rest of the band: only refinements of what is already non-zero
            for (; k <= se; k++) {
                int16_t &z = b[kZigzag[k]];
                if (z != 0 && br.bit() && (z & p1) == 0) z = (int16_t)(z >= 0 ? z + p1 : z + m1);
            }
            eobrun--;
        }
    }

    size_t parse_scan(size_t at, size_t len) { // `at`: first byte after the segment length; returns the offset after the entropy-coded data
        const int ns = file[at];
        if (ns < 1 || ns > (int)comps.size() || len < (size_t)(4 + 2 * ns)) throw Error("jpeg: bad scan header");
        std::vector<Component *> sc;
        for (int i = 0; i < ns; i++) {
            Component *c = nullptr;
            for (Component &q : comps)
                if (q.id == file[at + 1 + 2 * i]) c = &q;
            if (!c) throw Error("jpeg: scan refers to an unknown component");
            c->td = file[at + 2 + 2 * i] >> 4, c->ta = file[at + 2 + 2 * i] & 15;
            if (c->td > 3 || c->ta > 3) throw Error("jpeg: bad table selector");
            sc.push_back(c);
        }
        const int ss = file[at + 1 + 2 * ns], se = file[at + 2 + 2 * ns], ah = file[at + 3 + 2 * ns] >> 4, al = file[at + 3 + 2 * ns] & 15;
        if (progressive) {
            if (ss > se || se > 63 || (ss == 0 && se != 0) || (ss > 0 && ns != 1) || al > 13) throw Error("jpeg: bad progressive scan parameters");
        } else if (ss != 0 || se != 63 || ah != 0 || al != 0) throw Error("jpeg: bad sequential scan parameters");
        for (Component *c : sc) {
            if ((!progressive || (ss == 0 && ah == 0)) && !dc[c->td].set) throw Error("jpeg: missing DC Huffman table");
            if ((!progressive || ss > 0) && !ac[c->ta].set) throw Error("jpeg: missing AC Huffman table");
        }
        BitReader br(&file[at + len], file.data() + file.size());
        eobrun = 0;
        for (Component *c : sc) c->dc_pred = 0;
        auto one = [&](Component &c, int bx, int by) {
            int16_t *b = &c.coef[((size_t)by * c.bw + bx) * 64];
            if (!progressive) block_sequential(br, c, b);
            else if (ss == 0) ah == 0 ? block_dc_first(br, c, b, al) : block_dc_refine(br, b, al);
            else ah == 0 ? block_ac_first(br, c, b, ss, se, al) : block_ac_refine(br, c, b, ss, se, al);
        };
        const bool interleaved = ns > 1;
        const int units_x = interleaved ? mcux : sc[0]->cw, units_y = interleaved ? mcuy : sc[0]->ch;
        int until_restart = restart_interval, next_rst = 0;
        for (int uy = 0; uy < units_y; uy++)
            for (int ux = 0; ux < units_x; ux++) {
                if (restart_interval && until_restart == 0) { // RSTn: byte-align, skip the marker, reset the predictions
                    const uint8_t *p = br.p;
                    while (p + 1 < br.end && !(p[0] == 0xFF && p[1] >= 0xD0 && p[1] <= 0xD7)) p++;
                    if (p + 1 >= br.end || p[1] != 0xD0 + next_rst) throw Error("jpeg: restart marker missing or out of order");
                    br.p = p + 2;
                    br.reset();
                    next_rst = (next_rst + 1) & 7;
                    until_restart = restart_interval;
                    eobrun = 0;
                    for (Component *c : sc) c->dc_pred = 0;
                }
                if (interleaved) {
                    for (Component *c : sc)
                        for (int v = 0; v < c->v; v++)
                            for (int h = 0; h < c->h; h++) one(*c, ux * c->h + h, uy * c->v + v);
                } else one(*sc[0], ux, uy);
                until_restart--;
            }
        // the next marker
        const uint8_t *p = br.p;
        while (p + 1 < br.end && !(p[0] == 0xFF && p[1] != 0x00 && !(p[1] >= 0xD0 && p[1] <= 0xD7))) p++;
        return (size_t)(p - file.data());
    }

    void decode(const std::string &path) {
        if (file.size() < 4 || file[0] != 0xFF || file[1] != 0xD8) throw Error("not a JPEG file: " + path);
        size_t pos = 2;
        bool eoi = false;
        while (!eoi) {
            while (pos < file.size() && file[pos] != 0xFF) pos++; // (garbage between segments is skipped)
            while (pos < file.size() && file[pos] == 0xFF) pos++;
            if (pos >= file.size()) break; // a missing EOI is tolerated once the scans are in
            const int m = file[pos++];
            if (m == 0xD9) {
                eoi = true;
                break;
            }
            if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
            const size_t len = be16(pos);
            if (len < 2 || pos + len > file.size()) throw Error("jpeg: truncated segment in " + path);
            const size_t at = pos + 2, n = len - 2;
            if (m == 0xDB) parse_dqt(at, n);
            else if (m == 0xC4) parse_dht(at, n);
            else if (m == 0xDD) restart_interval = be16(at);
            else if (m >= 0xC0 && m <= 0xCF && m != 0xC8 && m != 0xCC) parse_sof(at, m);
            else if (m == 0xDA) {
                if (!have_sof) throw Error("jpeg: scan before the frame header");
                pos = parse_scan(at, n);
                continue;
            }
            pos += len;
        }
        if (!have_sof) throw Error("jpeg: no frame header in " + path);
    }

    // planes of samples (component resolution, padded to whole blocks)
    std::vector<uint8_t> plane(const Component &c) const {
        if (!qt_set[c.tq]) throw Error("jpeg: missing quantisation table");
        std::vector<uint8_t> out((size_t)c.bw * 8 * c.bh * 8);
        for (int by = 0; by < c.bh; by++)
            for (int bx = 0; bx < c.bw; bx++) idct_islow(&c.coef[((size_t)by * c.bw + bx) * 64], qt[c.tq], &out[((size_t)by * 8) * c.bw * 8 + (size_t)bx * 8], c.bw * 8);
        return out;
    }
};

// IJG jdsample.c: h2v1 / h2v2 "fancy" (triangle filter) upsampling and plain replication for the other combinations; `sw`, `sh` = the
// component's own size in samples (edge samples are replicated), output width x height.
std::vector<uint8_t> upsample(const std::vector<uint8_t> &in, int stride, int sw, int sh, int h, int v, int hmax, int vmax, int width, int height) {
    std::vector<uint8_t> out((size_t)width * height);
    const int fx = hmax / h, fy = vmax / v;
    auto at = [&](int x, int y) { return (int)in[(size_t)std::min(std::max(y, 0), sh - 1) * stride + std::min(std::max(x, 0), sw - 1)]; };
    if (fx == 1 && fy == 1) {
        for (int y = 0; y < height; y++)
            for (int x = 0; x < width; x++) out[(size_t)y * width + x] = (uint8_t)at(x, y);
    } else if (fx == 2 && fy == 1) { // h2v1_fancy: 3/4 nearer + 1/4 further; the first and last column are copied
        for (int y = 0; y < height; y++)
            for (int x = 0; x < width; x++) {
                const int i = x >> 1, cur = at(i, y);
                int r;
                if (sw == 1) r = cur;
                else if (x == 0) r = cur;
                else if (i == sw - 1 && (x & 1)) r = cur;
                else r = (x & 1) ? (3 * cur + at(i + 1, y) + 2) >> 2 : (3 * cur + at(i - 1, y) + 1) >> 2;
                out[(size_t)y * width + x] = (uint8_t)r;
            }
    } else if (fx == 2 && fy == 2) { // h2v2_fancy: 9/16, 3/16, 3/16, 1/16
        for (int y = 0; y < height; y++) {
            const int j = y >> 1, jn = (y & 1) ? j + 1 : j - 1; // the nearer neighbouring row (clamped by at())
            for (int x = 0; x < width; x++) {
                const int i = x >> 1;
                const int thiscol = 3 * at(i, j) + at(i, jn);
                int r;
                if (sw == 1) r = (thiscol * 4 + 8) >> 4;
                else if (x == 0) r = (thiscol * 4 + 8) >> 4;
                else if (i == sw - 1 && (x & 1)) r = (thiscol * 4 + 7) >> 4;
                else if (x & 1) r = (thiscol * 3 + (3 * at(i + 1, j) + at(i + 1, jn)) + 7) >> 4;
                else r = (thiscol * 3 + (3 * at(i - 1, j) + at(i - 1, jn)) + 8) >> 4;
                out[(size_t)y * width + x] = (uint8_t)r;
            }
        }
    } else if (fx == 1 && fy == 2) { // h1v2_fancy (libjpeg-turbo): 3/4 this row + 1/4 the nearer neighbouring row, bias 1 (upper) / 2 (lower)
        for (int y = 0; y < height; y++) {
            const int j = y >> 1, jn = (y & 1) ? j + 1 : j - 1;
            for (int x = 0; x < width; x++) out[(size_t)y * width + x] = (uint8_t)((3 * at(x, j) + at(x, jn) + ((y & 1) ? 2 : 1)) >> 2);
        }
    } else { // anything else: replication (jdsample.c int_upsample)
        for (int y = 0; y < height; y++)
            for (int x = 0; x < width; x++) out[(size_t)y * width + x] = (uint8_t)at(x / fx, y / fy);
    }
    return out;
}

} // namespace

Bitmap Bitmap::read_jpeg(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Error("cannot open " + path);
    Decoder d;
    d.file.assign((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    d.decode(path);
    const int W = d.width, H = d.height;
    std::vector<std::vector<uint8_t>> full;
    for (const Component &c : d.comps) {
        const std::vector<uint8_t> p = d.plane(c);
        const int sw = (W * c.h + d.hmax - 1) / d.hmax, sh = (H * c.v + d.vmax - 1) / d.vmax;
        full.push_back(upsample(p, c.bw * 8, sw, sh, c.h, c.v, d.hmax, d.vmax, W, H));
    }
    Bitmap b;
    b.size_x = (uint32_t)W, b.size_y = (uint32_t)H;
    b.colors.resize((size_t)3 * W * H);
    auto clamp = [](int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); };
    for (size_t i = 0; i < (size_t)W * H; i++) {
        int r, g, bl;
        if (full.size() == 1) r = g = bl = full[0][i];
        else { // IJG jdcolor.c: 16-bit fixed point, ONE_HALF rounding
            const int y = full[0][i], cb = full[1][i] - 128, cr = full[2][i] - 128;
            r = clamp(y + (int)((91881L * cr + 32768) >> 16));
            g = clamp(y + (int)((-22554L * cb - 46802L * cr + 32768) >> 16));
            bl = clamp(y + (int)((116130L * cb + 32768) >> 16));
        }
        b.colors[3 * i] = (float)r / 255.0f, b.colors[3 * i + 1] = (float)g / 255.0f, b.colors[3 * i + 2] = (float)bl / 255.0f;
    }
    return b;
}

} // namespace rlh
