// rt_jpeg.cpp -- baseline JPEG decoder for the image-texture path (host only).
//
// The reference loads its one texture asset, earthmap.jpg, through RtwImage (reference RtwImage.h:51-87), which calls
// stbi_loadf from the vendored third-party stb_image v2.30 (reference StbImageImpl.cpp:19-21, external/stb_image.h).
// ImageTexture then looks up NEAREST texels (Texture.h:110-133), so parity with the reference needs the decoded
// bytes themselves, not just a picture that looks alike: this file restates the arithmetic stb_image v2.30 publishes
// for baseline JPEG so that the texels come out byte for byte the same --
//   * Huffman decoding per ITU T.81 F.2.2 (any conforming decoder agrees on the coefficients),
//   * the integer inverse DCT of the IJG "slow-but-accurate" kind with stb's 12-bit constants, its column pass
//     rounded at >> 10 and its row pass at >> 17 with the +128 level shift folded in,
//   * chroma upsampling by stb's separable (3,1)/4 tent filters ("h_2", "v_2", "hv_2"),
//   * stb's fixed-point YCbCr -> RGB (20 fractional bits, the Cb term of green masked to 16 bits),
// and nothing else of that library (no progressive mode, no other file formats: RT_ERR_UNSUPPORTED).
// tests/test_texture_pipeline.py checks the result against the texels the reference's own stb path produced
// (tests/golden/earthmap_rgb8.npz) and, where oracle/_ref is built, against stb itself on subsampled, greyscale and
// restart-interval files.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rt_scenes_c.h"

void rt_set_error(const char* fmt, ...); // rt_error.cpp

namespace {

struct JpegError {
    int status;
    std::string what;
};
[[noreturn]] void Fail(int status, const char* what) { throw JpegError{status, what}; }

// T.81 Figure A.6
const uint8_t kZigzag[64 + 15] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,
                                  6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31,
                                  39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
                                  // a corrupt run may step past 63: land somewhere harmless
                                  63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};

struct Huffman {
    // canonical code tables, T.81 Annex C: codes of length L are consecutive starting at first[L]
    int maxcode[18];  // largest code of length L, left-aligned to 16 bits, +1
    int delta[17];    // index of the first symbol of length L minus its code
    uint8_t symbols[256];
    uint8_t sizes[257];
    uint16_t codes[256];
    bool present = false;
    void Build(const int* counts)
    {
        int k = 0;
        for (int len = 1; len <= 16; ++len)
            for (int c = 0; c < counts[len - 1]; ++c) {
                if (k >= 256) Fail(RT_ERR_INVALID, "corrupt JPEG: Huffman table with more than 256 codes");
                sizes[k++] = (uint8_t)len;
            }
        sizes[k] = 0;
        int code = 0;
        k = 0;
        for (int len = 1; len <= 16; ++len) {
            delta[len] = k - code;
            if (sizes[k] == len) {
                while (sizes[k] == len) codes[k++] = (uint16_t)code++;
                if (code - 1 >= (1 << len)) Fail(RT_ERR_INVALID, "corrupt JPEG: bad Huffman code lengths");
            }
            maxcode[len] = code << (16 - len);
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        present = true;
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int dcPred = 0;
    int x = 0, y = 0;   // size in samples
    int w2 = 0, h2 = 0; // size padded to whole MCUs
    std::vector<uint8_t> data;
};

struct Decoder {
    const uint8_t* p;
    const uint8_t* end;
    uint16_t dequant[4][64];
    bool dequantPresent[4] = {false, false, false, false};
    Huffman dc[4], ac[4];
    Component comp[3];
    int nComp = 0, width = 0, height = 0, hMax = 1, vMax = 1;
    int restartInterval = 0;
    bool jfif = false;
    int adobeTransform = -1, rgbIds = 0;
    // entropy-coded segment reader
    uint32_t bitBuf = 0;
    int bitCount = 0;
    int marker = 0; // marker met inside the entropy-coded data (0 = none)
    bool noMore = false;

    int Get8()
    {
        if (p >= end) Fail(RT_ERR_INVALID, "corrupt JPEG: unexpected end of data");
        return *p++;
    }
    int Get16()
    {
        const int hi = Get8();
        return (hi << 8) | Get8();
    }

    void GrowBits()
    {
        do {
            int b = noMore ? 0 : (p < end ? *p++ : 0);
            if (b == 0xff && !noMore) {
                int c = p < end ? *p++ : 0;
                while (c == 0xff) c = p < end ? *p++ : 0; // fill bytes
                if (c != 0) { // a marker ends the segment; feed zeros from here on
                    marker = c;
                    noMore = true;
                    b = 0;
                }
            }
            bitBuf |= (uint32_t)b << (24 - bitCount);
            bitCount += 8;
        } while (bitCount <= 24);
    }

    int DecodeHuff(const Huffman& h)
    {
        if (bitCount < 16) GrowBits();
        const int top = (int)(bitBuf >> 16);
        int len = 1;
        while (top >= h.maxcode[len]) ++len;
        if (len == 17 || len > bitCount) Fail(RT_ERR_INVALID, "corrupt JPEG: bad Huffman code");
        const int idx = (int)(bitBuf >> (32 - len)) + h.delta[len];
        if (idx < 0 || idx >= 256) Fail(RT_ERR_INVALID, "corrupt JPEG: bad Huffman code");
        bitBuf <<= len;
        bitCount -= len;
        return h.symbols[idx];
    }

    // T.81 F.2.2.1 RECEIVE + EXTEND
    int ReceiveExtend(int n)
    {
        if (n == 0) return 0;
        if (bitCount < n) GrowBits();
        const int v = (int)(bitBuf >> (32 - n));
        bitBuf <<= n;
        bitCount -= n;
        return v < (1 << (n - 1)) ? v - (1 << n) + 1 : v;
    }

    void DecodeBlock(short* data, Component& c)
    {
        std::memset(data, 0, 64 * sizeof(short));
        const Huffman& hd = dc[c.td];
        const Huffman& ha = ac[c.ta];
        const uint16_t* dq = dequant[c.tq];
        const int t = DecodeHuff(hd);
        if (t > 15) Fail(RT_ERR_INVALID, "corrupt JPEG: bad DC category");
        c.dcPred += ReceiveExtend(t);
        data[0] = (short)(c.dcPred * dq[0]);
        int k = 1;
        do {
            const int rs = DecodeHuff(ha);
            const int s = rs & 15, r = rs >> 4;
            if (s == 0) {
                if (rs != 0xf0) break; // end of block
                k += 16;
            } else {
                k += r;
                const int zig = kZigzag[k++];
                data[zig] = (short)(ReceiveExtend(s) * dq[zig]);
            }
        } while (k < 64);
    }

    void ResetEntropy()
    {
        bitBuf = 0;
        bitCount = 0;
        noMore = false;
        marker = 0;
        for (int k = 0; k < nComp; ++k) comp[k].dcPred = 0;
    }
};

inline uint8_t Clamp8(int x) { return (unsigned)x > 255u ? (x < 0 ? 0 : 255) : (uint8_t)x; }

// stb_image v2.30 stbi__idct_block: constants are round(x * 4096).
#define RT_F2F(x) ((int)((x)*4096 + 0.5))
#define RT_FSH(x) ((x)*4096)
#define RT_IDCT_1D(s0, s1, s2, s3, s4, s5, s6, s7)      \
    int t0, t1, t2, t3, p1, p2, p3, p4, p5, x0, x1, x2, x3; \
    p2 = s2;                                            \
    p3 = s6;                                            \
    p1 = (p2 + p3) * RT_F2F(0.5411961f);                \
    t2 = p1 + p3 * RT_F2F(-1.847759065f);               \
    t3 = p1 + p2 * RT_F2F(0.765366865f);                \
    p2 = s0;                                            \
    p3 = s4;                                            \
    t0 = RT_FSH(p2 + p3);                               \
    t1 = RT_FSH(p2 - p3);                               \
    x0 = t0 + t3;                                       \
    x3 = t0 - t3;                                       \
    x1 = t1 + t2;                                       \
    x2 = t1 - t2;                                       \
    t0 = s7;                                            \
    t1 = s5;                                            \
    t2 = s3;                                            \
    t3 = s1;                                            \
    p3 = t0 + t2;                                       \
    p4 = t1 + t3;                                       \
    p1 = t0 + t3;                                       \
    p2 = t1 + t2;                                       \
    p5 = (p3 + p4) * RT_F2F(1.175875602f);              \
    t0 = t0 * RT_F2F(0.298631336f);                     \
    t1 = t1 * RT_F2F(2.053119869f);                     \
    t2 = t2 * RT_F2F(3.072711026f);                     \
    t3 = t3 * RT_F2F(1.501321110f);                     \
    p1 = p5 + p1 * RT_F2F(-0.899976223f);               \
    p2 = p5 + p2 * RT_F2F(-2.562915447f);               \
    p3 = p3 * RT_F2F(-1.961570560f);                    \
    p4 = p4 * RT_F2F(-0.390180644f);                    \
    t3 += p1 + p4;                                      \
    t2 += p2 + p3;                                      \
    t1 += p2 + p4;                                      \
    t0 += p1 + p3;

void IdctBlock(uint8_t* out, int outStride, const short* data)
{
    int val[64];
    int* v = val;
    const short* d = data;
    for (int i = 0; i < 8; ++i, ++d, ++v) {
        if (d[8] == 0 && d[16] == 0 && d[24] == 0 && d[32] == 0 && d[40] == 0 && d[48] == 0 && d[56] == 0) {
            const int dcterm = d[0] * 4;
            v[0] = v[8] = v[16] = v[24] = v[32] = v[40] = v[48] = v[56] = dcterm;
        } else {
            RT_IDCT_1D(d[0], d[8], d[16], d[24], d[32], d[40], d[48], d[56])
            x0 += 512;
            x1 += 512;
            x2 += 512;
            x3 += 512;
            v[0] = (x0 + t3) >> 10;
            v[56] = (x0 - t3) >> 10;
            v[8] = (x1 + t2) >> 10;
            v[48] = (x1 - t2) >> 10;
            v[16] = (x2 + t1) >> 10;
            v[40] = (x2 - t1) >> 10;
            v[24] = (x3 + t0) >> 10;
            v[32] = (x3 - t0) >> 10;
        }
    }
    v = val;
    uint8_t* o = out;
    for (int i = 0; i < 8; ++i, v += 8, o += outStride) {
        RT_IDCT_1D(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7])
        x0 += 65536 + (128 << 17);
        x1 += 65536 + (128 << 17);
        x2 += 65536 + (128 << 17);
        x3 += 65536 + (128 << 17);
        o[0] = Clamp8((x0 + t3) >> 17);
        o[7] = Clamp8((x0 - t3) >> 17);
        o[1] = Clamp8((x1 + t2) >> 17);
        o[6] = Clamp8((x1 - t2) >> 17);
        o[2] = Clamp8((x2 + t1) >> 17);
        o[5] = Clamp8((x2 - t1) >> 17);
        o[3] = Clamp8((x3 + t0) >> 17);
        o[4] = Clamp8((x3 - t0) >> 17);
    }
}

// ---- chroma upsampling, one output row at a time (near = the closer source row, far = the other one)
inline uint8_t Div4(int x) { return (uint8_t)(x >> 2); }
inline uint8_t Div16(int x) { return (uint8_t)(x >> 4); }

const uint8_t* Resample1(uint8_t*, const uint8_t* nearRow, const uint8_t*, int, int) { return nearRow; }

const uint8_t* ResampleV2(uint8_t* out, const uint8_t* nearRow, const uint8_t* farRow, int w, int)
{
    for (int i = 0; i < w; ++i) out[i] = Div4(3 * nearRow[i] + farRow[i] + 2);
    return out;
}

const uint8_t* ResampleH2(uint8_t* out, const uint8_t* in, const uint8_t*, int w, int)
{
    if (w == 1) {
        out[0] = out[1] = in[0];
        return out;
    }
    out[0] = in[0];
    out[1] = Div4(in[0] * 3 + in[1] + 2);
    int i;
    for (i = 1; i < w - 1; ++i) {
        const int n = 3 * in[i] + 2;
        out[i * 2 + 0] = Div4(n + in[i - 1]);
        out[i * 2 + 1] = Div4(n + in[i + 1]);
    }
    out[i * 2 + 0] = Div4(in[w - 2] * 3 + in[w - 1] + 2);
    out[i * 2 + 1] = in[w - 1];
    return out;
}

const uint8_t* ResampleHV2(uint8_t* out, const uint8_t* nearRow, const uint8_t* farRow, int w, int)
{
    if (w == 1) {
        out[0] = out[1] = Div4(3 * nearRow[0] + farRow[0] + 2);
        return out;
    }
    int t1 = 3 * nearRow[0] + farRow[0];
    out[0] = Div4(t1 + 2);
    for (int i = 1; i < w; ++i) {
        const int t0 = t1;
        t1 = 3 * nearRow[i] + farRow[i];
        out[i * 2 - 1] = Div16(3 * t0 + t1 + 8);
        out[i * 2] = Div16(3 * t1 + t0 + 8);
    }
    out[w * 2 - 1] = Div4(t1 + 2);
    return out;
}

const uint8_t* ResampleGeneric(uint8_t* out, const uint8_t* nearRow, const uint8_t*, int w, int hs)
{
    for (int i = 0; i < w; ++i)
        for (int j = 0; j < hs; ++j) out[i * hs + j] = nearRow[i];
    return out;
}

using ResampleFn = const uint8_t* (*)(uint8_t*, const uint8_t*, const uint8_t*, int, int);

// stb_image v2.30 stbi__YCbCr_to_RGB_row
#define RT_FLOAT2FIXED(x) (((int)((x)*4096.0f + 0.5f)) << 8)
void YCbCrToRgbRow(uint8_t* out, const uint8_t* y, const uint8_t* pcb, const uint8_t* pcr, int count)
{
    for (int i = 0; i < count; ++i) {
        const int yFixed = (y[i] << 20) + (1 << 19); // rounding
        const int cr = pcr[i] - 128, cb = pcb[i] - 128;
        int r = yFixed + cr * RT_FLOAT2FIXED(1.40200f);
        int g = yFixed + (cr * -RT_FLOAT2FIXED(0.71414f)) + ((cb * -RT_FLOAT2FIXED(0.34414f)) & 0xffff0000);
        int b = yFixed + cb * RT_FLOAT2FIXED(1.77200f);
        r >>= 20;
        g >>= 20;
        b >>= 20;
        out[0] = Clamp8(r);
        out[1] = Clamp8(g);
        out[2] = Clamp8(b);
        out += 3;
    }
}

void ParseTables(Decoder& z, int m)
{
    int len = z.Get16() - 2;
    if (len < 0) Fail(RT_ERR_INVALID, "corrupt JPEG: bad segment length");
    switch (m) {
    case 0xDB: // DQT
        while (len > 0) {
            const int q = z.Get8();
            const int prec = q >> 4, t = q & 15;
            if ((prec != 0 && prec != 1) || t > 3) Fail(RT_ERR_INVALID, "corrupt JPEG: bad DQT");
            for (int i = 0; i < 64; ++i) z.dequant[t][kZigzag[i]] = (uint16_t)(prec ? z.Get16() : z.Get8());
            z.dequantPresent[t] = true;
            len -= prec ? 129 : 65;
        }
        if (len != 0) Fail(RT_ERR_INVALID, "corrupt JPEG: bad DQT length");
        return;
    case 0xC4: // DHT
        while (len > 0) {
            const int q = z.Get8();
            const int tc = q >> 4, th = q & 15;
            if (tc > 1 || th > 3) Fail(RT_ERR_INVALID, "corrupt JPEG: bad DHT header");
            int counts[16], n = 0;
            for (int i = 0; i < 16; ++i) {
                counts[i] = z.Get8();
                n += counts[i];
            }
            if (n > 256) Fail(RT_ERR_INVALID, "corrupt JPEG: bad DHT counts");
            Huffman& h = tc == 0 ? z.dc[th] : z.ac[th];
            for (int i = 0; i < n; ++i) h.symbols[i] = (uint8_t)z.Get8();
            h.Build(counts);
            len -= 17 + n;
        }
        if (len != 0) Fail(RT_ERR_INVALID, "corrupt JPEG: bad DHT length");
        return;
    case 0xDD: // DRI
        if (len != 2) Fail(RT_ERR_INVALID, "corrupt JPEG: bad DRI length");
        z.restartInterval = z.Get16();
        return;
    case 0xE0: { // APP0: a "JFIF" tag says the three components are YCbCr whatever an Adobe segment claims
        static const char tag[5] = {'J', 'F', 'I', 'F', '\0'};
        if (len >= 5) {
            bool ok = true;
            for (int i = 0; i < 5; ++i)
                if (z.Get8() != (uint8_t)tag[i]) ok = false;
            len -= 5;
            if (ok) z.jfif = true;
        }
        break;
    }
    case 0xEE: { // APP14 "Adobe": colour transform flag
        static const char tag[6] = {'A', 'd', 'o', 'b', 'e', '\0'};
        if (len >= 12) {
            bool ok = true;
            for (int i = 0; i < 6; ++i)
                if (z.Get8() != (uint8_t)tag[i]) ok = false;
            len -= 6;
            if (ok) {
                z.Get8();  // version
                z.Get16(); // flags0
                z.Get16(); // flags1
                z.adobeTransform = z.Get8();
                len -= 6;
            }
        }
        break;
    }
    default:
        if (!((m >= 0xE0 && m <= 0xEF) || m == 0xFE)) Fail(RT_ERR_UNSUPPORTED, "JPEG: unsupported marker"); // APPn, COM
        break;
    }
    if (z.end - z.p < len) Fail(RT_ERR_INVALID, "corrupt JPEG: bad segment length");
    z.p += len; // the rest of an APPn / COM segment
}

int NextMarker(Decoder& z)
{
    int x = z.Get8();
    while (x != 0xff) x = z.Get8(); // (garbage between segments is skipped)
    while (x == 0xff) x = z.Get8();
    return x;
}

void ParseFrame(Decoder& z)
{
    const int len = z.Get16();
    const int prec = z.Get8();
    if (prec != 8) Fail(RT_ERR_UNSUPPORTED, "JPEG: only 8-bit samples are supported");
    z.height = z.Get16();
    z.width = z.Get16();
    if (z.height <= 0 || z.width <= 0) Fail(RT_ERR_INVALID, "corrupt JPEG: zero image size");
    z.nComp = z.Get8();
    if (z.nComp != 1 && z.nComp != 3) Fail(RT_ERR_UNSUPPORTED, "JPEG: only greyscale and 3-component images are supported");
    if (len != 8 + 3 * z.nComp) Fail(RT_ERR_INVALID, "corrupt JPEG: bad SOF length");
    static const char rgbIds[3] = {'R', 'G', 'B'};
    for (int i = 0; i < z.nComp; ++i) {
        Component& c = z.comp[i];
        c.id = z.Get8();
        if (z.nComp == 3 && c.id == rgbIds[i]) ++z.rgbIds;
        const int q = z.Get8();
        c.h = q >> 4;
        c.v = q & 15;
        if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4) Fail(RT_ERR_INVALID, "corrupt JPEG: bad sampling factors");
        c.tq = z.Get8();
        if (c.tq > 3) Fail(RT_ERR_INVALID, "corrupt JPEG: bad quantisation table index");
        z.hMax = std::max(z.hMax, c.h);
        z.vMax = std::max(z.vMax, c.v);
    }
    for (int i = 0; i < z.nComp; ++i)
        if (z.hMax % z.comp[i].h != 0 || z.vMax % z.comp[i].v != 0) Fail(RT_ERR_UNSUPPORTED, "JPEG: fractional sampling ratios");
    if ((long long)z.width * z.height > (1ll << 28)) Fail(RT_ERR_UNSUPPORTED, "JPEG: image too large");
    const int mcuW = z.hMax * 8, mcuH = z.vMax * 8;
    const int mcuX = (z.width + mcuW - 1) / mcuW, mcuY = (z.height + mcuH - 1) / mcuH;
    for (int i = 0; i < z.nComp; ++i) {
        Component& c = z.comp[i];
        c.x = (z.width * c.h + z.hMax - 1) / z.hMax;
        c.y = (z.height * c.v + z.vMax - 1) / z.vMax;
        c.w2 = mcuX * c.h * 8;
        c.h2 = mcuY * c.v * 8;
        c.data.assign((size_t)c.w2 * c.h2, 0);
    }
}

void ParseScanHeader(Decoder& z, int* order, int& scanN)
{
    const int len = z.Get16();
    scanN = z.Get8();
    if (scanN < 1 || scanN > 3 || scanN > z.nComp) Fail(RT_ERR_INVALID, "corrupt JPEG: bad SOS component count");
    if (len != 6 + 2 * scanN) Fail(RT_ERR_INVALID, "corrupt JPEG: bad SOS length");
    for (int i = 0; i < scanN; ++i) {
        const int id = z.Get8(), q = z.Get8();
        int which = 0;
        while (which < z.nComp && z.comp[which].id != id) ++which;
        if (which == z.nComp) Fail(RT_ERR_INVALID, "corrupt JPEG: SOS names an unknown component");
        z.comp[which].td = q >> 4;
        z.comp[which].ta = q & 15;
        if (z.comp[which].td > 3 || z.comp[which].ta > 3) Fail(RT_ERR_INVALID, "corrupt JPEG: bad Huffman table index");
        order[i] = which;
    }
    const int ss = z.Get8();
    z.Get8(); // Se
    const int a = z.Get8();
    if (ss != 0 || a != 0) Fail(RT_ERR_INVALID, "corrupt JPEG: bad baseline SOS parameters");
}

void DecodeScan(Decoder& z, const int* order, int scanN)
{
    for (int i = 0; i < scanN; ++i) {
        const Component& c = z.comp[order[i]];
        if (!z.dc[c.td].present || !z.ac[c.ta].present || !z.dequantPresent[c.tq])
            Fail(RT_ERR_INVALID, "corrupt JPEG: scan uses a table that was never defined");
    }
    z.ResetEntropy();
    short block[64];
    int todo = z.restartInterval ? z.restartInterval : 0x7fffffff;
    auto restart = [&]() {
        if (z.bitCount < 24) z.GrowBits();
        if (z.marker < 0xD0 || z.marker > 0xD7) return false; // not a restart marker: the data is over
        z.ResetEntropy();
        todo = z.restartInterval;
        return true;
    };
    if (scanN == 1) {
        // non-interleaved: the component's own blocks, row by row
        Component& c = z.comp[order[0]];
        const int bw = (c.x + 7) >> 3, bh = (c.y + 7) >> 3;
        for (int j = 0; j < bh; ++j)
            for (int i = 0; i < bw; ++i) {
                z.DecodeBlock(block, c);
                IdctBlock(c.data.data() + (size_t)c.w2 * j * 8 + i * 8, c.w2, block);
                if (--todo <= 0 && !restart()) return;
            }
        return;
    }
    const int mcuW = z.hMax * 8, mcuH = z.vMax * 8;
    const int mcuX = (z.width + mcuW - 1) / mcuW, mcuY = (z.height + mcuH - 1) / mcuH;
    for (int j = 0; j < mcuY; ++j)
        for (int i = 0; i < mcuX; ++i) {
            for (int k = 0; k < scanN; ++k) {
                Component& c = z.comp[order[k]];
                for (int y = 0; y < c.v; ++y)
                    for (int x = 0; x < c.h; ++x) {
                        const int x2 = (i * c.h + x) * 8, y2 = (j * c.v + y) * 8;
                        z.DecodeBlock(block, c);
                        IdctBlock(c.data.data() + (size_t)c.w2 * y2 + x2, c.w2, block);
                    }
            }
            if (--todo <= 0 && !restart()) return;
        }
}

// -> interleaved RGB8, row 0 = top
void DecodeJpeg(const uint8_t* bytes, size_t n, int& width, int& height, std::vector<uint8_t>* rgb)
{
    Decoder z{};
    z.p = bytes;
    z.end = bytes + n;
    if (n < 4 || z.Get8() != 0xff || z.Get8() != 0xD8) Fail(RT_ERR_INVALID, "not a JPEG file (no SOI marker)");
    int m = NextMarker(z);
    while (!(m == 0xC0 || m == 0xC1)) {
        if (m == 0xC2) Fail(RT_ERR_UNSUPPORTED, "progressive JPEG is not supported (baseline only: re-save the texture as baseline)");
        if (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) Fail(RT_ERR_UNSUPPORTED, "JPEG: unsupported frame type");
        ParseTables(z, m);
        m = NextMarker(z);
    }
    ParseFrame(z);
    width = z.width;
    height = z.height;
    if (!rgb) return;
    bool sawScan = false;
    m = NextMarker(z);
    while (m != 0xD9) { // EOI
        if (m == 0xDA) {
            int order[3], scanN = 0;
            ParseScanHeader(z, order, scanN);
            DecodeScan(z, order, scanN);
            sawScan = true;
            if (z.marker) { // the scan ran into a marker: that is the next segment
                m = z.marker;
                z.marker = 0;
                if (m >= 0xD0 && m <= 0xD7) m = NextMarker(z);
                continue;
            }
        } else if (m == 0xDC) { // DNL
            z.Get16();
            z.Get16();
        } else {
            ParseTables(z, m);
        }
        if (z.p >= z.end) break; // (no EOI: accept what was decoded, as stb does)
        m = NextMarker(z);
    }
    if (!sawScan) Fail(RT_ERR_INVALID, "corrupt JPEG: no scan");

    // upsample + colour convert, one output row at a time (stb_image v2.30 load_jpeg_image)
    struct Resampler {
        ResampleFn fn;
        const uint8_t *line0, *line1;
        int hs, vs, wLores, ystep, ypos;
        std::vector<uint8_t> buf;
    } rs[3];
    for (int k = 0; k < z.nComp; ++k) {
        Resampler& r = rs[k];
        r.hs = z.hMax / z.comp[k].h;
        r.vs = z.vMax / z.comp[k].v;
        r.ystep = r.vs >> 1;
        r.wLores = (z.width + r.hs - 1) / r.hs;
        r.ypos = 0;
        r.line0 = r.line1 = z.comp[k].data.data();
        r.buf.assign((size_t)z.width + 3 + 8, 0);
        if (r.hs == 1 && r.vs == 1) r.fn = Resample1;
        else if (r.hs == 1 && r.vs == 2) r.fn = ResampleV2;
        else if (r.hs == 2 && r.vs == 1) r.fn = ResampleH2;
        else if (r.hs == 2 && r.vs == 2) r.fn = ResampleHV2;
        else r.fn = ResampleGeneric;
    }
    rgb->assign((size_t)z.width * z.height * 3, 0);
    for (int j = 0; j < z.height; ++j) {
        const uint8_t* rows[3] = {nullptr, nullptr, nullptr};
        for (int k = 0; k < z.nComp; ++k) {
            Resampler& r = rs[k];
            const bool yBot = r.ystep >= (r.vs >> 1);
            rows[k] = r.fn(r.buf.data(), yBot ? r.line1 : r.line0, yBot ? r.line0 : r.line1, r.wLores, r.hs);
            if (++r.ystep >= r.vs) {
                r.ystep = 0;
                r.line0 = r.line1;
                if (++r.ypos < z.comp[k].y) r.line1 += z.comp[k].w2;
            }
        }
        uint8_t* out = rgb->data() + (size_t)j * z.width * 3;
        if (z.nComp == 3) {
            if (z.rgbIds == 3 || (z.adobeTransform == 0 && !z.jfif)) { // stored as RGB, not YCbCr
                for (int i = 0; i < z.width; ++i) {
                    out[3 * i] = rows[0][i];
                    out[3 * i + 1] = rows[1][i];
                    out[3 * i + 2] = rows[2][i];
                }
            } else {
                YCbCrToRgbRow(out, rows[0], rows[1], rows[2], z.width);
            }
        } else {
            for (int i = 0; i < z.width; ++i) out[3 * i] = out[3 * i + 1] = out[3 * i + 2] = rows[0][i];
        }
    }
}

} // namespace

extern "C" {

int rt_image_decode_jpeg(const uint8_t* bytes, uint64_t n_bytes, int32_t* width, int32_t* height, uint8_t* rgb_out,
                         uint64_t capacity, int32_t linearize)
{
    if (!bytes || !width || !height) {
        rt_set_error("rt_image_decode_jpeg: NULL argument");
        return RT_ERR_INVALID;
    }
    try {
        int w = 0, h = 0;
        if (!rgb_out) {
            DecodeJpeg(bytes, (size_t)n_bytes, w, h, nullptr);
            *width = w;
            *height = h;
            return RT_OK;
        }
        std::vector<uint8_t> rgb;
        DecodeJpeg(bytes, (size_t)n_bytes, w, h, &rgb);
        *width = w;
        *height = h;
        if ((uint64_t)rgb.size() > capacity) {
            rt_set_error("rt_image_decode_jpeg: the image needs %zu bytes, the buffer holds %llu", rgb.size(), (unsigned long long)capacity);
            return RT_ERR_INVALID;
        }
        if (linearize)
            rt_image_linearize_rgb8(rgb.data(), rgb_out, rgb.size());
        else
            std::memcpy(rgb_out, rgb.data(), rgb.size());
        return RT_OK;
    } catch (const JpegError& e) {
        rt_set_error("rt_image_decode_jpeg: %s", e.what.c_str());
        return e.status;
    } catch (const std::exception& e) {
        rt_set_error("rt_image_decode_jpeg: %s", e.what());
        return RT_ERR_INVALID;
    }
}

int rt_image_load(const char* path, int32_t* width, int32_t* height, uint8_t* rgb_out, uint64_t capacity)
{
    if (!path || !width || !height) {
        rt_set_error("rt_image_load: NULL argument");
        return RT_ERR_INVALID;
    }
    FILE* f = std::fopen(path, "rb");
    if (!f) {
        rt_set_error("rt_image_load: cannot open %s", path);
        return RT_ERR_INVALID;
    }
    std::vector<uint8_t> bytes;
    uint8_t chunk[65536];
    size_t got;
    while ((got = std::fread(chunk, 1, sizeof chunk, f)) > 0) bytes.insert(bytes.end(), chunk, chunk + got);
    std::fclose(f);
    return rt_image_decode_jpeg(bytes.data(), bytes.size(), width, height, rgb_out, capacity, 1);
}

} // extern "C"
