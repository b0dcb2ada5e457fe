// measured.cpp -- host-side load of an irregularly sampled isotropic BRDF
// (scenes/brdfs/*.brdf) into the kd-tree the GPU's IrregIsotropicBRDF::f walks.
//
// Mirrors, step for step (so every float matches the reference's tables):
//   ReadFloatFile                     core/floatfile.cpp:38-83
//   MeasuredMaterial ctor (.brdf)     materials/measured.cpp:85-135
//   RGBSpectrum::FromSampled          core/spectrum.h:428-451
//   InterpolateSpectrumSamples        core/spectrum.cpp:207-220
//   BRDFRemap                         core/reflection.cpp:247-256
//   KdTree ctor / recursiveBuild      core/kdtree.h:101-147
// sinf/cosf/atan2f here are the host libm's, as in the reference.
#include "scene.h"

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>

namespace pb {

namespace {

const float kCIE[471][3] = {
#include "cie1931.inc"
};

bool ReadFloatFile(const std::string &filename, std::vector<float> *values) {
    FILE *f = fopen(filename.c_str(), "r");
    if (!f) return false;
    int c;
    bool inNumber = false;
    char cur[64];
    int pos = 0;
    while ((c = getc(f)) != EOF) {
        if (inNumber) {
            if (isdigit(c) || c == '.' || c == 'e' || c == '-' || c == '+') {
                if (pos < 62) cur[pos++] = (char)c;
            } else {
                cur[pos] = '\0';
                values->push_back((float)atof(cur));
                inNumber = false;
                pos = 0;
            }
        } else {
            if (isdigit(c) || c == '.' || c == '-' || c == '+') {
                inNumber = true;
                cur[pos++] = (char)c;
            } else if (c == '#') {
                while ((c = getc(f)) != '\n' && c != EOF) {}
            }
        }
    }
    fclose(f);
    return true;
}

float InterpolateSpectrumSamples(const float *lambda, const float *vals, int n, float l) {
    if (l <= lambda[0]) return vals[0];
    if (l >= lambda[n - 1]) return vals[n - 1];
    for (int i = 0; i < n - 1; ++i)
        if (l >= lambda[i] && l <= lambda[i + 1]) {
            float t = (l - lambda[i]) / (lambda[i + 1] - lambda[i]);
            return Lerp(t, vals[i], vals[i + 1]);
        }
    return 0.f;
}

// RGBSpectrum::FromSampled for already-sorted wavelengths
void SampledToRGB(const float *lambda, const float *v, int n, float rgb[3]) {
    float xyz[3] = {0, 0, 0};
    float yint = 0.f;
    for (int i = 0; i < 471; ++i) {
        yint += kCIE[i][1];
        float val = InterpolateSpectrumSamples(lambda, v, n, float(360 + i));
        xyz[0] += val * kCIE[i][0];
        xyz[1] += val * kCIE[i][1];
        xyz[2] += val * kCIE[i][2];
    }
    xyz[0] /= yint;
    xyz[1] /= yint;
    xyz[2] /= yint;
    // XYZToRGB, core/spectrum.h:51-55
    rgb[0] = 3.240479f * xyz[0] - 1.537150f * xyz[1] - 0.498535f * xyz[2];
    rgb[1] = -0.969256f * xyz[0] + 1.875991f * xyz[1] + 0.041556f * xyz[2];
    rgb[2] = 0.055648f * xyz[0] - 0.204043f * xyz[1] + 1.057311f * xyz[2];
}

inline float SinTheta(const V3 &w) { return sqrtf(std::max(0.f, 1.f - w.z * w.z)); }   // reflection.h:66-71
inline float SphericalPhi(const V3 &v) {   // geometry.h:644-647
    float p = atan2f(v.y, v.x);
    return (p < 0.f) ? p + 2.f * kPi : p;
}
V3 BRDFRemap(const V3 &wo, const V3 &wi) {   // reflection.cpp:247-256
    float cosi = wi.z, coso = wo.z;
    float sini = SinTheta(wi), sino = SinTheta(wo);
    float phii = SphericalPhi(wi), phio = SphericalPhi(wo);
    float dphi = phii - phio;
    if (dphi < 0.) dphi += 2.f * kPi;
    if (dphi > 2.f * kPi) dphi -= 2.f * kPi;
    if (dphi > kPi) dphi = 2.f * kPi - dphi;
    return V3(sini * sino, dphi / kPi, cosi * coso);
}
inline V3 SphericalDirection(float sintheta, float costheta, float phi) {   // geometry.h:626-630
    return V3(sintheta * cosf(phi), sintheta * sinf(phi), costheta);
}

struct KdBuild {
    const std::vector<PbrtGpuBrdfSample> &src;
    std::vector<PbrtGpuKdNode> &nodes;
    std::vector<PbrtGpuBrdfSample> &data;
    uint32_t nextFreeNode = 1;
    static uint32_t Pack(uint32_t axis, uint32_t hasLeft, uint32_t right) {
        return (axis & 3u) | ((hasLeft & 1u) << 2) | (right << 3);
    }
    void Build(uint32_t nodeNum, int start, int end, std::vector<uint32_t> &order) {
        if (start + 1 == end) {   // KdNode::initLeaf
            nodes[nodeNum].splitPos = 0.f;
            nodes[nodeNum].bits = Pack(3, 0, (1u << 29) - 1);
            data[nodeNum] = src[order[start]];
            return;
        }
        BBox bound;
        for (int i = start; i < end; ++i) {
            const float *p = src[order[i]].p;
            bound = Union(bound, V3(p[0], p[1], p[2]));
        }
        int splitAxis = bound.MaximumExtent();
        int splitPos = (start + end) / 2;
        // CompareNode breaks ties on the ADDRESS of the sample in the input
        // vector (kdtree.h:93-96) == its index
        std::nth_element(order.begin() + start, order.begin() + splitPos, order.begin() + end,
                         [&](uint32_t a, uint32_t b) {
                             float pa = src[a].p[splitAxis], pb_ = src[b].p[splitAxis];
                             return pa == pb_ ? (a < b) : pa < pb_;
                         });
        uint32_t hasLeft = 0, right = (1u << 29) - 1;
        nodes[nodeNum].splitPos = src[order[splitPos]].p[splitAxis];
        data[nodeNum] = src[order[splitPos]];
        if (start < splitPos) {
            hasLeft = 1;
            uint32_t child = nextFreeNode++;
            nodes[nodeNum].bits = Pack(splitAxis, hasLeft, right);
            Build(child, start, splitPos, order);
        }
        if (splitPos + 1 < end) {
            right = nextFreeNode++;
            nodes[nodeNum].bits = Pack(splitAxis, hasLeft, right);
            Build(right, splitPos + 1, end, order);
        }
        nodes[nodeNum].bits = Pack(splitAxis, hasLeft, right);
    }
};

}  // namespace

// (for the "spectrum" parameter type, parser.cpp)
bool ReadSpectrumFile(const std::string &filename, std::vector<float> *values) { return ReadFloatFile(filename, values); }
void SortedSamplesToRGB(const float *lambda, const float *v, int n, float rgb[3]) { SampledToRGB(lambda, v, n, rgb); }

bool LoadIrregIsotropicBRDF(const std::string &filename, MeasuredData *out, std::string *err) {
    std::vector<float> values;
    if (!ReadFloatFile(filename, &values) || values.empty()) {
        *err = "Unable to read BRDF data from file \"" + filename + "\"";
        return false;
    }
    uint32_t pos = 0;
    int numWls = int(values[pos++]);
    if (numWls <= 0 || (values.size() - 1 - numWls) % (4 + numWls) != 0) {
        *err = "Excess or insufficient data in theta, phi BRDF file \"" + filename + "\"";
        return false;
    }
    std::vector<float> wls;
    for (int i = 0; i < numWls; ++i) wls.push_back(values[pos++]);
    for (int i = 0; i + 1 < numWls; ++i)
        if (wls[i] > wls[i + 1]) { *err = "unsorted wavelengths in \"" + filename + "\" are not supported"; return false; }
    std::vector<PbrtGpuBrdfSample> samples;
    while (pos < values.size()) {
        float thetai = values[pos++];
        float phii = values[pos++];
        float thetao = values[pos++];
        float phio = values[pos++];
        V3 wo = SphericalDirection(sinf(thetao), cosf(thetao), phio);
        V3 wi = SphericalDirection(sinf(thetai), cosf(thetai), phii);
        PbrtGpuBrdfSample s;
        SampledToRGB(wls.data(), &values[pos], numWls, s.v);
        pos += numWls;
        V3 p = BRDFRemap(wo, wi);
        s.p[0] = p.x; s.p[1] = p.y; s.p[2] = p.z;
        samples.push_back(s);
    }
    uint32_t n = (uint32_t)samples.size();
    out->nodes.assign(n, PbrtGpuKdNode());
    out->data.assign(n, PbrtGpuBrdfSample());
    std::vector<uint32_t> order(n);
    for (uint32_t i = 0; i < n; ++i) order[i] = i;
    KdBuild kb{samples, out->nodes, out->data};
    kb.Build(0, 0, (int)n, order);
    return true;
}

}  // namespace pb
