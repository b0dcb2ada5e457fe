// envmap.cpp -- what InfiniteAreaLight's constructor computes (lights/infinite.cpp:67-104), on the host, once per light:
// the image file -> texels x (L * scale) -> MIPMap<RGBSpectrum> (core/mipmap.h:112-200: Lanczos resampling to a power of
// two, then the box-filtered pyramid) -> img[] = luminance of the map filtered at 1 / max(w, h), times sin(theta) ->
// Distribution2D (core/montecarlo.h:54-80, montecarlo.cpp:358-370).  Level 0 of the pyramid and the distribution's arrays
// cross the ABI as a PbrtGpuEnvMap; the kernels and the oracle only ever do MIPMap::Lookup(s, t) at width 0 (the bilinear
// `triangle` over level 0) and Distribution2D::SampleContinuous / Pdf.
//
// Every float operation below is in the reference's order and type (its M_PI is a float literal, core/pbrt.h:188-191, so
// nothing here is promoted to double); libm's sinf / logf are the ones the reference binary links.  tests/test_builder_pins.py compares the
// arrays with the reference's own members (oracle/tools/refdump.cpp `envmap`).
#include "scene.h"

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>

namespace pb {
namespace {

constexpr float PI_REF = 3.14159265358979323846f;   // the reference's M_PI

struct Rgb {
    float c[3];
    Rgb() { c[0] = c[1] = c[2] = 0.f; }
    explicit Rgb(float v) { c[0] = c[1] = c[2] = v; }
};
inline Rgb operator+(const Rgb &a, const Rgb &b) { Rgb r; for (int i = 0; i < 3; ++i) r.c[i] = a.c[i] + b.c[i]; return r; }
inline Rgb operator*(float f, const Rgb &a) { Rgb r; for (int i = 0; i < 3; ++i) r.c[i] = a.c[i] * f; return r; }
inline float Luminance(const Rgb &a) {   // RGBSpectrum::y(), core/spectrum.h:424-427
    const float YWeight[3] = {0.212671f, 0.715160f, 0.072169f};
    return YWeight[0] * a.c[0] + YWeight[1] * a.c[1] + YWeight[2] * a.c[2];
}
inline int Floor2IntH(float v) { return (int)floorf(v); }
inline int ModH(int a, int b) {   // core/pbrt.h:237-242
    int n = int(a / b);
    a -= n * b;
    if (a < 0) a += b;
    return a;
}
inline float Log2H(float x) {   // core/pbrt.h:255-258
    static float invLog2 = 1.f / logf(2.f);
    return logf(x) * invLog2;
}
inline bool IsPowerOf2H(int v) { return v && !(v & (v - 1)); }
inline uint32_t RoundUpPow2H(uint32_t v) {
    v--;
    v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
    return v + 1;
}
float LanczosH(float x, float tau = 2) {   // core/texture.cpp:266-274
    x = fabsf(x);
    if (x < 1e-5) return 1;
    if (x > 1.) return 0;
    x *= PI_REF;
    float s = sinf(x * tau) / (x * tau);
    float lanczos = sinf(x) / x;
    return s * lanczos;
}
struct ResampleWeight {
    int firstTexel;
    float weight[4];
};
std::vector<ResampleWeight> ResampleWeights(uint32_t oldres, uint32_t newres) {   // core/mipmap.h:67-86
    std::vector<ResampleWeight> wt(newres);
    float filterwidth = 2.f;
    for (uint32_t i = 0; i < newres; ++i) {
        float center = (i + .5f) * oldres / newres;
        wt[i].firstTexel = Floor2IntH((center - filterwidth) + 0.5f);
        for (int j = 0; j < 4; ++j) {
            float pos = wt[i].firstTexel + j + .5f;
            wt[i].weight[j] = LanczosH((pos - center) / filterwidth);
        }
        float invSumWts = 1.f / (wt[i].weight[0] + wt[i].weight[1] + wt[i].weight[2] + wt[i].weight[3]);
        for (uint32_t j = 0; j < 4; ++j) wt[i].weight[j] *= invSumWts;
    }
    return wt;
}
inline float ClampLow(float v) {   // Clamp(v, 0, INFINITY), core/pbrt.h: NaN falls through both comparisons
    if (v < 0.f) return 0.f;
    if (v > INFINITY) return INFINITY;
    return v;
}

// MIPMap<RGBSpectrum>(xres, yres, data) with the defaults InfiniteAreaLight passes: doTri false, TEXTURE_REPEAT
struct MipMap {
    struct Level {
        uint32_t us, vs;
        std::vector<Rgb> a;
    };
    std::vector<Level> pyramid;
    uint32_t width, height, nLevels;

    const Rgb &Texel(uint32_t level, int s, int t) const {   // mipmap.h:203-228, TEXTURE_REPEAT
        const Level &l = pyramid[level];
        s = ModH(s, (int)l.us);
        t = ModH(t, (int)l.vs);
        return l.a[(size_t)t * l.us + s];
    }
    MipMap(uint32_t sres, uint32_t tres, const Rgb *img) {
        std::vector<Rgb> resampled;
        if (!IsPowerOf2H((int)sres) || !IsPowerOf2H((int)tres)) {
            uint32_t sPow2 = RoundUpPow2H(sres), tPow2 = RoundUpPow2H(tres);
            std::vector<ResampleWeight> sWeights = ResampleWeights(sres, sPow2);
            resampled.assign((size_t)sPow2 * tPow2, Rgb());
            for (uint32_t t = 0; t < tres; ++t)
                for (uint32_t s = 0; s < sPow2; ++s) {
                    Rgb acc(0.f);
                    for (int j = 0; j < 4; ++j) {
                        int origS = ModH(sWeights[s].firstTexel + j, (int)sres);
                        if (origS >= 0 && origS < (int)sres) acc = acc + sWeights[s].weight[j] * img[(size_t)t * sres + origS];
                    }
                    resampled[(size_t)t * sPow2 + s] = acc;
                }
            std::vector<ResampleWeight> tWeights = ResampleWeights(tres, tPow2);
            std::vector<Rgb> workData(tPow2);
            for (uint32_t s = 0; s < sPow2; ++s) {
                for (uint32_t t = 0; t < tPow2; ++t) {
                    Rgb acc(0.f);
                    for (uint32_t j = 0; j < 4; ++j) {
                        int offset = ModH(tWeights[t].firstTexel + (int)j, (int)tres);
                        if (offset >= 0 && offset < (int)tres) acc = acc + tWeights[t].weight[j] * resampled[(size_t)offset * sPow2 + s];
                    }
                    workData[t] = acc;
                }
                for (uint32_t t = 0; t < tPow2; ++t)
                    for (int k = 0; k < 3; ++k) resampled[(size_t)t * sPow2 + s].c[k] = ClampLow(workData[t].c[k]);
            }
            img = resampled.data();
            sres = sPow2;
            tres = tPow2;
        }
        width = sres;
        height = tres;
        nLevels = 1 + Floor2IntH(Log2H(float(std::max(sres, tres))));
        pyramid.resize(nLevels);
        pyramid[0].us = sres; pyramid[0].vs = tres;
        pyramid[0].a.assign(img, img + (size_t)sres * tres);
        for (uint32_t i = 1; i < nLevels; ++i) {
            uint32_t sRes = std::max(1u, pyramid[i - 1].us / 2);
            uint32_t tRes = std::max(1u, pyramid[i - 1].vs / 2);
            pyramid[i].us = sRes; pyramid[i].vs = tRes;
            pyramid[i].a.resize((size_t)sRes * tRes);
            for (uint32_t t = 0; t < tRes; ++t)
                for (uint32_t s = 0; s < sRes; ++s)
                    pyramid[i].a[(size_t)t * sRes + s] =
                        .25f * (Texel(i - 1, 2 * s, 2 * t) + Texel(i - 1, 2 * s + 1, 2 * t) + Texel(i - 1, 2 * s, 2 * t + 1) +
                                Texel(i - 1, 2 * s + 1, 2 * t + 1));
        }
    }
    Rgb Triangle(uint32_t level, float s, float t) const {   // mipmap.h:258-269
        level = std::min(level, nLevels - 1);
        s = s * pyramid[level].us - 0.5f;
        t = t * pyramid[level].vs - 0.5f;
        int s0 = Floor2IntH(s), t0 = Floor2IntH(t);
        float ds = s - s0, dt = t - t0;
        return (1.f - ds) * (1.f - dt) * Texel(level, s0, t0) + (1.f - ds) * dt * Texel(level, s0, t0 + 1) +
               ds * (1.f - dt) * Texel(level, s0 + 1, t0) + ds * dt * Texel(level, s0 + 1, t0 + 1);
    }
    Rgb Lookup(float s, float t, float w) const {   // mipmap.h:238-255
        float level = nLevels - 1 + Log2H(std::max(w, 1e-8f));
        if (level < 0) return Triangle(0, s, t);
        else if (level >= nLevels - 1) return Texel(nLevels - 1, 0, 0);
        else {
            uint32_t iLevel = (uint32_t)Floor2IntH(level);
            float delta = level - iLevel;
            return (1.f - delta) * Triangle(iLevel, s, t) + delta * Triangle(iLevel + 1, s, t);
        }
    }
};

// Distribution1D(f, n), montecarlo.h:56-80: func copied, cdf (n + 1 entries) and funcInt
float Distribution1D(const float *f, int n, float *cdf) {
    cdf[0] = 0.;
    for (int i = 1; i < n + 1; ++i) cdf[i] = cdf[i - 1] + f[i - 1] / n;
    float funcInt = cdf[n];
    if (funcInt == 0.f) {
        for (int i = 1; i < n + 1; ++i) cdf[i] = float(i) / float(n);
    } else {
        for (int i = 1; i < n + 1; ++i) cdf[i] /= funcInt;
    }
    return funcInt;
}

inline bool IsWhitespace(char c) { return c == ' ' || c == '\n' || c == '\t'; }   // imageio.cpp:288-290
// one header word and the single whitespace after it (imageio.cpp:302-322)
int ReadWord(FILE *fp, char *buffer, int bufferLength) {
    int n = 0;
    int c = fgetc(fp);
    while (c != EOF && !IsWhitespace((char)c) && n < bufferLength) {
        buffer[n++] = (char)c;
        c = fgetc(fp);
    }
    if (n < bufferLength) { buffer[n] = '\0'; return n; }
    return -1;
}

}  // namespace

// ReadImagePFM, core/imageio.cpp:330-408: "PF" (RGB) or "Pf" (grey), rows in file order, |scale| applied, the sign of
// scale says little- (negative) or big-endian
bool ReadImagePFM(const std::string &filename, int *xres, int *yres, std::vector<float> *rgb, std::string *err) {
    FILE *fp = fopen(filename.c_str(), "rb");
    auto fail = [&]() {
        *err = "Error reading PFM file \"" + filename + "\"";
        if (fp) fclose(fp);
        return false;
    };
    if (!fp) return fail();
    char buffer[80];
    int nChannels;
    if (ReadWord(fp, buffer, 80) == -1) return fail();
    if (strcmp(buffer, "Pf") == 0) nChannels = 1;
    else if (strcmp(buffer, "PF") == 0) nChannels = 3;
    else return fail();
    if (ReadWord(fp, buffer, 80) == -1) return fail();
    const int width = atoi(buffer);
    if (ReadWord(fp, buffer, 80) == -1) return fail();
    const int height = atoi(buffer);
    if (ReadWord(fp, buffer, 80) == -1) return fail();
    float scale = 0.f;
    if (sscanf(buffer, "%f", &scale) != 1) return fail();
    if (width <= 0 || height <= 0 || (long long)width * height > (1ll << 28)) return fail();
    const size_t nFloats = (size_t)nChannels * width * height;
    {   // (the data must be there before memory is set aside for it)
        const long at = ftell(fp);
        fseek(fp, 0, SEEK_END);
        const long end = ftell(fp);
        fseek(fp, at, SEEK_SET);
        if (at < 0 || end < at || (size_t)(end - at) < nFloats * sizeof(float)) return fail();
    }
    std::vector<float> data(nFloats);
    if (fread(data.data(), sizeof(float), nFloats, fp) != nFloats) return fail();
    const bool fileLittleEndian = scale < 0.f;
    const uint32_t one = 1;
    const bool hostLittleEndian = *reinterpret_cast<const uint8_t *>(&one) == 1;
    if (hostLittleEndian ^ fileLittleEndian)
        for (size_t i = 0; i < nFloats; ++i) {
            uint8_t b[4];
            memcpy(b, &data[i], 4);
            std::swap(b[0], b[3]);
            std::swap(b[1], b[2]);
            memcpy(&data[i], b, 4);
        }
    if (fabsf(scale) != 1.f)
        for (size_t i = 0; i < nFloats; ++i) data[i] *= fabsf(scale);
    rgb->resize((size_t)3 * width * height);
    for (size_t i = 0; i < (size_t)width * height; ++i)
        for (int k = 0; k < 3; ++k) (*rgb)[3 * i + k] = nChannels == 1 ? data[i] : data[3 * i + k];
    *xres = width;
    *yres = height;
    fclose(fp);
    return true;
}

bool ReadImage(const std::string &filename, int *xres, int *yres, std::vector<float> *rgb, std::string *err) {
    const std::string suffix = filename.size() >= 5 ? filename.substr(filename.size() - 4) : "";
    if (suffix == ".pfm" || suffix == ".PFM") return ReadImagePFM(filename, xres, yres, rgb, err);
    if (suffix == ".exr" || suffix == ".EXR") return ReadImageEXR(filename, xres, yres, rgb, err);
    if (suffix == ".tga" || suffix == ".TGA") return ReadImageTGA(filename, xres, yres, rgb, err);
    *err = "image files are read from .pfm, .tga and .exr only: " + filename;
    return false;
}

// InfiniteAreaLight::InfiniteAreaLight, lights/infinite.cpp:67-104.  `rgb` NULL: no map, one texel = L
void BuildEnvMap(const float L[3], int width, int height, const float *rgb, EnvMap *out) {
    std::vector<Rgb> texels((size_t)width * height);
    for (size_t i = 0; i < texels.size(); ++i)
        for (int k = 0; k < 3; ++k) texels[i].c[k] = rgb ? rgb[3 * i + k] * L[k] : L[k];
    MipMap radianceMap((uint32_t)width, (uint32_t)height, texels.data());
    float filter = 1.f / std::max(width, height);
    std::vector<float> img((size_t)width * height);
    for (int v = 0; v < height; ++v) {
        float vp = (float)v / (float)height;
        float sinTheta = sinf(PI_REF * float(v + .5f) / float(height));
        for (int u = 0; u < width; ++u) {
            float up = (float)u / (float)width;
            img[u + (size_t)v * width] = Luminance(radianceMap.Lookup(up, vp, filter));
            img[u + (size_t)v * width] *= sinTheta;
        }
    }
    out->width = (int)radianceMap.width;
    out->height = (int)radianceMap.height;
    out->texels.resize((size_t)3 * out->width * out->height);
    for (size_t i = 0; i < radianceMap.pyramid[0].a.size(); ++i)
        for (int k = 0; k < 3; ++k) out->texels[3 * i + k] = radianceMap.pyramid[0].a[i].c[k];
    // Distribution2D(img, width, height), montecarlo.cpp:358-370
    out->nu = width;
    out->nv = height;
    out->func = img;
    out->cdf.resize((size_t)height * (width + 1));
    out->funcInt.resize(height);
    for (int v = 0; v < height; ++v)
        out->funcInt[v] = Distribution1D(&img[(size_t)v * width], width, &out->cdf[(size_t)v * (width + 1)]);
    out->margCdf.resize(height + 1);
    out->margFuncInt = Distribution1D(out->funcInt.data(), height, out->margCdf.data());
}

}  // namespace pb
