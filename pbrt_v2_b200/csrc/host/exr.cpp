// exr.cpp -- a reader for the OpenEXR files environment maps come in (scenes/textures/*.exr of the reference are scan-line
// files with HALF R, G, B channels, ZIP-compressed), standing in for ReadImageEXR (core/imageio.cpp:146-167), which hands the
// job to the OpenEXR library's RgbaInputFile.  Written from the file format's layout; no OpenEXR code here or in the image.
//
// What RgbaInputFile does and this reader keeps: the data window's pixels, rows by increasing y whatever the file's
// lineOrder, R / G / B as halfs converted (exactly) to float, other channels skipped.  What it does not do: tiled, deep
// or multi-part files, luminance/chroma files, subsampled channels, UINT / FLOAT colour channels (RgbaInputFile would
// round those to half), PIZ / PXR24 / B44 / DWA compression -- all refused with an error naming the feature.
// Parity note: the reference build this repo pins against has no OpenEXR, so this decoder is
// checked against an independent reader and writer in tests/test_exr.py, not against the reference binary.
#include "scene.h"

#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <map>

namespace pb {
namespace {

float HalfToFloat(uint16_t h) {
    const uint32_t sign = (uint32_t)(h >> 15) << 31;
    uint32_t exp = (h >> 10) & 0x1f, man = h & 0x3ff;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else {   // a subnormal half is a normal float
            int e = -1;
            do { ++e; man <<= 1; } while (!(man & 0x400));
            bits = sign | (uint32_t)(127 - 15 - e) << 23 | (man & 0x3ff) << 13;
        }
    } else if (exp == 31) bits = sign | 0x7f800000u | man << 13;
    else bits = sign | (exp + 127 - 15) << 23 | man << 13;
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

struct Channel {
    std::string name;
    int type;        // 0 UINT, 1 HALF, 2 FLOAT
    int bytes;
};

template <typename T> bool Get(const std::vector<uint8_t> &d, size_t &p, T *v) {
    if (p + sizeof(T) > d.size()) return false;
    memcpy(v, d.data() + p, sizeof(T));
    p += sizeof(T);
    return true;
}
bool GetString(const std::vector<uint8_t> &d, size_t &p, std::string *s) {
    const size_t b = p;
    while (p < d.size() && d[p] != 0) ++p;
    if (p >= d.size() || p - b > 255) return false;
    s->assign((const char *)d.data() + b, p - b);
    ++p;
    return true;
}

// the two steps every byte-oriented EXR compressor shares: undo the delta predictor, then re-interleave the two halves
void Unpredict(std::vector<uint8_t> &t, std::vector<uint8_t> *out) {
    for (size_t i = 1; i < t.size(); ++i) t[i] = (uint8_t)(t[i - 1] + t[i] - 128);
    out->resize(t.size());
    const size_t half = (t.size() + 1) / 2;
    for (size_t i = 0, a = 0, b = half; i < t.size();) {
        (*out)[i++] = t[a++];
        if (i < t.size()) (*out)[i++] = t[b++];
    }
}
bool UnRle(const uint8_t *in, size_t n, size_t expect, std::vector<uint8_t> *out) {
    out->clear();
    size_t p = 0;
    while (p < n) {
        const int8_t c = (int8_t)in[p++];
        if (c < 0) {   // -c literal bytes
            const size_t k = (size_t)(-(int)c);
            if (p + k > n) return false;
            out->insert(out->end(), in + p, in + p + k);
            p += k;
        } else {       // c + 1 copies of the next byte
            if (p >= n) return false;
            out->insert(out->end(), (size_t)c + 1, in[p++]);
        }
        if (out->size() > expect) return false;
    }
    return out->size() == expect;
}

}  // namespace

bool ReadImageEXR(const std::string &filename, int *xres, int *yres, std::vector<float> *rgb, std::string *err) {
    auto fail = [&](const std::string &why) {
        *err = "Unable to read image file \"" + filename + "\": " + why;
        return false;
    };
    std::vector<uint8_t> d;
    {
        FILE *f = fopen(filename.c_str(), "rb");
        if (!f) return fail("cannot open");
        fseek(f, 0, SEEK_END);
        const long n = ftell(f);
        fseek(f, 0, SEEK_SET);
        if (n < 0 || n > (1l << 31)) { fclose(f); return fail("file too large"); }
        d.resize((size_t)n);
        const bool ok = fread(d.data(), 1, d.size(), f) == d.size();
        fclose(f);
        if (!ok) return fail("short read");
    }
    size_t p = 0;
    uint32_t magic = 0, version = 0;
    if (!Get(d, p, &magic) || !Get(d, p, &version) || magic != 20000630u) return fail("not an OpenEXR file");
    if ((version & 0xff) != 2) return fail("unknown file version");
    if (version & 0x200) return fail("tiled files are not supported");
    if (version & (0x800 | 0x1000)) return fail("deep / multi-part files are not supported");
    std::vector<Channel> channels;
    int compression = -1, lineOrder = 0;
    int32_t dw[4] = {0, 0, -1, -1};
    bool haveDw = false;
    for (;;) {   // attributes: name\0 type\0 int32 size, value; an empty name ends the header
        std::string name, type;
        if (!GetString(d, p, &name)) return fail("truncated header");
        if (name.empty()) break;
        int32_t size = 0;
        if (!GetString(d, p, &type) || !Get(d, p, &size) || size < 0 || p + (size_t)size > d.size()) return fail("truncated header");
        size_t q = p;
        p += (size_t)size;
        if (name == "channels") {
            for (;;) {
                Channel c;
                if (!GetString(d, q, &c.name)) return fail("bad channel list");
                if (c.name.empty()) break;
                int32_t pixelType, xs, ys;
                uint8_t lin[4];
                if (!Get(d, q, &pixelType) || !Get(d, q, &lin) || !Get(d, q, &xs) || !Get(d, q, &ys)) return fail("bad channel list");
                if (pixelType < 0 || pixelType > 2) return fail("unknown channel type");
                if (xs != 1 || ys != 1) return fail("subsampled channels (luminance/chroma files) are not supported");
                c.type = pixelType;
                c.bytes = pixelType == 1 ? 2 : 4;
                channels.push_back(c);
            }
        } else if (name == "compression") {
            if (size != 1) return fail("bad compression attribute");
            compression = d[q];
        } else if (name == "dataWindow") {
            if (size != 16) return fail("bad dataWindow attribute");
            memcpy(dw, d.data() + q, 16);
            haveDw = true;
        } else if (name == "lineOrder") {
            if (size != 1) return fail("bad lineOrder attribute");
            lineOrder = d[q];
        }
    }
    (void)lineOrder;   // chunks carry their own y: any order is read the same way
    if (channels.empty() || compression < 0 || !haveDw) return fail("missing required attributes");
    const long long width = (long long)dw[2] - dw[0] + 1, height = (long long)dw[3] - dw[1] + 1;
    if (width <= 0 || height <= 0 || width * height > (1ll << 28)) return fail("bad data window");
    int linesPerBlock;
    switch (compression) {
    case 0: case 1: case 2: linesPerBlock = 1; break;   // none, RLE, ZIPS
    case 3: linesPerBlock = 16; break;                  // ZIP
    default: return fail("compression method " + std::to_string(compression) + " (PIZ / PXR24 / B44 / DWA) is not supported");
    }
    // the colour channels and where they lie inside a scan line (channels are stored in the list's -- alphabetical -- order)
    size_t lineBytes = 0;
    std::map<std::string, size_t> offsetOf;
    for (const Channel &c : channels) {
        if ((c.name == "R" || c.name == "G" || c.name == "B") && c.type != 1)
            return fail("colour channels other than HALF are not supported");
        offsetOf[c.name] = lineBytes;
        lineBytes += (size_t)c.bytes * (size_t)width;
    }
    if (!offsetOf.count("R") && !offsetOf.count("G") && !offsetOf.count("B")) return fail("no R, G or B channel (luminance-only files are not supported)");
    // (deflate expands at most ~1032 : 1, the RLE here 128 : 1: a data window the file cannot possibly fill is a corrupt header)
    if ((double)lineBytes * (double)height > 1100.0 * (double)d.size()) return fail("data window larger than the file can hold");
    const size_t nChunks = (size_t)((height + linesPerBlock - 1) / linesPerBlock);
    if (p + 8 * nChunks > d.size()) return fail("truncated offset table");
    rgb->assign((size_t)3 * width * height, 0.f);   // a missing colour channel reads as 0 (RgbaInputFile's fill value)
    std::vector<uint8_t> raw, tmp;
    std::vector<bool> seen((size_t)height, false);
    for (size_t c = 0; c < nChunks; ++c) {
        uint64_t off;
        memcpy(&off, d.data() + p + 8 * c, 8);
        size_t q = (size_t)off;
        int32_t y, dataSize;
        if (off > d.size() || !Get(d, q, &y) || !Get(d, q, &dataSize) || dataSize < 0 || q + (size_t)dataSize > d.size())
            return fail("bad chunk offset");
        if (y < dw[1] || y > dw[3] || (y - dw[1]) % linesPerBlock) return fail("bad chunk scan line");
        const int lines = (int)std::min<long long>(linesPerBlock, (long long)dw[3] - y + 1);
        const size_t expect = lineBytes * (size_t)lines;
        const uint8_t *src = d.data() + q;
        if ((size_t)dataSize == expect || compression == 0) {   // stored as is (compression that does not help is skipped)
            if ((size_t)dataSize != expect) return fail("bad chunk size");
            raw.assign(src, src + expect);
        } else if (compression == 1) {
            if (!UnRle(src, (size_t)dataSize, expect, &tmp)) return fail("bad RLE data");
            Unpredict(tmp, &raw);
        } else {
            tmp.resize(expect);
            uLongf n = (uLongf)expect;
            if (uncompress(tmp.data(), &n, src, (uLong)dataSize) != Z_OK || n != expect) return fail("bad ZIP data");
            Unpredict(tmp, &raw);
        }
        for (int l = 0; l < lines; ++l) {
            const size_t row = (size_t)(y - dw[1]) + (size_t)l;
            seen[row] = true;
            const uint8_t *line = raw.data() + lineBytes * (size_t)l;
            const char *names[3] = {"R", "G", "B"};
            for (int k = 0; k < 3; ++k) {
                auto it = offsetOf.find(names[k]);
                if (it == offsetOf.end()) continue;
                const uint8_t *ch = line + it->second;
                for (long long x = 0; x < width; ++x) {
                    uint16_t h;
                    memcpy(&h, ch + 2 * x, 2);
                    (*rgb)[3 * (row * (size_t)width + (size_t)x) + k] = HalfToFloat(h);
                }
            }
        }
    }
    for (bool s : seen)
        if (!s) return fail("missing scan lines");
    *xres = (int)width;
    *yres = (int)height;
    return true;
}


// WriteImageEXR (core/imageio.cpp:170-193): what RgbaOutputFile(name, displayWindow, dataWindow, WRITE_RGBA) stores -- HALF
// A, B, G, R channels over the data window (the crop region inside the full-resolution display window), alpha 1 --
// as a scan-line file with ZIP compression (the OpenEXR library's default would be PIZ; any reader takes either).
// float -> half as the library's `half` does it: round to nearest even, overflow to infinity.
static uint16_t FloatToHalf(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    const uint16_t sign = (uint16_t)((x >> 16) & 0x8000u);
    const uint32_t absx = x & 0x7fffffffu;
    if (absx >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | (absx > 0x7f800000u ? 0x200u | ((absx >> 13) & 0x3ffu) : 0u));   // inf / NaN
    if (absx >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);                 // rounds up to 65536 or beyond: infinity
    if (absx < 0x33000001u) return sign;                                      // at most half of the smallest subnormal: zero
    int e = (int)(absx >> 23) - 127;
    uint32_t m = (absx & 0x7fffffu) | 0x800000u;
    int shift;   // bits dropped from the 24-bit significand
    uint32_t base;
    if (e < -14) { shift = 13 + (-14 - e); base = 0; }                       // subnormal half
    else { shift = 13; base = (uint32_t)(e + 15) << 10; m &= 0x7fffffu; }
    const uint32_t kept = m >> shift, rem = m & ((1u << shift) - 1), halfway = 1u << (shift - 1);
    uint32_t h = base + kept;
    if (rem > halfway || (rem == halfway && (h & 1u))) ++h;                    // carries ripple into the exponent correctly
    return (uint16_t)(sign | h);
}

bool WriteImageEXR(const std::string &name, const std::vector<float> &rgb, int xRes, int yRes, int totalXRes, int totalYRes,
                   int xOffset, int yOffset) {
    std::vector<uint8_t> out;
    auto put = [&](const void *p, size_t n) { out.insert(out.end(), (const uint8_t *)p, (const uint8_t *)p + n); };
    auto put32 = [&](int32_t v) { put(&v, 4); };
    auto attr = [&](const char *nm, const char *type, const void *v, int32_t n) {
        put(nm, strlen(nm) + 1); put(type, strlen(type) + 1); put32(n); put(v, (size_t)n);
    };
    put32(20000630); put32(2);
    {
        std::vector<uint8_t> ch;
        for (const char *c : {"A", "B", "G", "R"}) {
            ch.push_back((uint8_t)c[0]); ch.push_back(0);
            const int32_t rec[4] = {1, 0, 1, 1};   // HALF, pLinear 0 (+ 3 reserved bytes), xSampling, ySampling
            ch.insert(ch.end(), (const uint8_t *)rec, (const uint8_t *)rec + 16);
        }
        ch.push_back(0);
        attr("channels", "chlist", ch.data(), (int32_t)ch.size());
    }
    const uint8_t zip = 3, increasingY = 0;
    attr("compression", "compression", &zip, 1);
    const int32_t dataWindow[4] = {xOffset, yOffset, xOffset + xRes - 1, yOffset + yRes - 1};
    const int32_t displayWindow[4] = {0, 0, totalXRes - 1, totalYRes - 1};
    attr("dataWindow", "box2i", dataWindow, 16);
    attr("displayWindow", "box2i", displayWindow, 16);
    attr("lineOrder", "lineOrder", &increasingY, 1);
    const float one = 1.f, centre[2] = {0.f, 0.f};
    attr("pixelAspectRatio", "float", &one, 4);
    attr("screenWindowCenter", "v2f", centre, 8);
    attr("screenWindowWidth", "float", &one, 4);
    out.push_back(0);
    const int nChunks = (yRes + 15) / 16;
    const size_t tableAt = out.size();
    out.resize(out.size() + 8 * (size_t)nChunks);
    const uint16_t alpha = FloatToHalf(1.f);
    std::vector<uint8_t> raw, tmp, comp;
    for (int c = 0; c < nChunks; ++c) {
        const int y0 = 16 * c, lines = std::min(16, yRes - y0);
        raw.clear();
        for (int y = y0; y < y0 + lines; ++y)
            for (int k = 0; k < 4; ++k)   // A, B, G, R: one channel after the other inside a scan line
                for (int x = 0; x < xRes; ++x) {
                    const uint16_t h = k == 0 ? alpha : FloatToHalf(rgb[3 * ((size_t)y * xRes + x) + (3 - k)]);
                    raw.push_back((uint8_t)(h & 0xff)); raw.push_back((uint8_t)(h >> 8));
                }
        tmp.resize(raw.size());   // even bytes first, then odd bytes; then the delta predictor
        for (size_t i = 0, a = 0, b = (raw.size() + 1) / 2; i < raw.size(); ++i) tmp[(i & 1) ? b++ : a++] = raw[i];
        for (size_t i = tmp.size(); i-- > 1;) tmp[i] = (uint8_t)(tmp[i] - tmp[i - 1] + 128);
        uLongf n = compressBound((uLong)tmp.size());
        comp.resize(n);
        const bool packed = compress(comp.data(), &n, tmp.data(), (uLong)tmp.size()) == Z_OK && n < raw.size();
        const uint64_t at = out.size();
        memcpy(out.data() + tableAt + 8 * (size_t)c, &at, 8);
        put32(yOffset + y0);
        put32((int32_t)(packed ? n : raw.size()));
        if (packed) put(comp.data(), n); else put(raw.data(), raw.size());
    }
    FILE *f = fopen(name.c_str(), "wb");
    if (!f) { fprintf(stderr, "Error: Unable to write image file \"%s\"\n", name.c_str()); return false; }
    const bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    fclose(f);
    return ok;
}

}  // namespace pb
