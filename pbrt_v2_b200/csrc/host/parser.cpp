// parser.cpp -- scene-description front end for the hot path's host side.
//
// pbrt-v2's own front end (core/pbrtlex.ll, pbrtparse.yy, api.cpp: ~8.7k lines)
// is out of scope; this is a small from-scratch reader for the directive subset
// the BASELINE.json config scenes use, with the same semantics where they affect
// the numbers that reach the GPU:
//   * numbers are (float)atof(token)                    core/pbrtlex.ll:165-168
//   * CTM handling, named coordinate systems, attribute stacks   core/api.cpp:700-905
//   * TransformCache: shapes share Transform objects    core/api.cpp:117-150
//   * Shape / Material / AreaLightSource / LightSource  core/api.cpp:956-1059
//   * light order = declaration order                   core/api.cpp:961-967,1055-1058
//   * MakeCamera / MakeSampler / MakeRenderer defaults  core/api.cpp:196-209,1210-1281
// Unsupported directives are reported as errors (never silently ignored).
#include "scene.h"

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <functional>
#include <iostream>
#include <set>
#include <sstream>

namespace pb {

// ----------------------------------------------------------------- ParamSet --
float ParamSet::FindOneFloat(const std::string &n, float d) const {
    auto it = floats.find(n);
    return (it != floats.end() && it->second.size() == 1) ? it->second[0] : d;
}
int ParamSet::FindOneInt(const std::string &n, int d) const {
    auto it = ints.find(n);
    return (it != ints.end() && it->second.size() == 1) ? it->second[0] : d;
}
bool ParamSet::FindOneBool(const std::string &n, bool d) const {
    auto it = bools.find(n);
    return (it != bools.end() && it->second.size() == 1) ? it->second[0] : d;
}
V3 ParamSet::FindOnePoint(const std::string &n, const V3 &d) const {
    auto it = points.find(n);
    return (it != points.end() && it->second.size() == 1) ? it->second[0] : d;
}
V3 ParamSet::FindOneSpectrum(const std::string &n, const V3 &d) const {
    auto it = spectra.find(n);
    return (it != spectra.end() && it->second.size() == 1) ? it->second[0] : d;
}
std::string ParamSet::FindOneString(const std::string &n, const std::string &d) const {
    auto it = strings.find(n);
    return (it != strings.end() && it->second.size() == 1) ? it->second[0] : d;
}
const std::vector<float> *ParamSet::FindFloat(const std::string &n) const {
    auto it = floats.find(n);
    return it == floats.end() ? nullptr : &it->second;
}
const std::vector<int> *ParamSet::FindInt(const std::string &n) const {
    auto it = ints.find(n);
    return it == ints.end() ? nullptr : &it->second;
}
const std::vector<V3> *ParamSet::FindPoint(const std::string &n) const {
    auto it = points.find(n);
    return it == points.end() ? nullptr : &it->second;
}
const std::vector<V3> *ParamSet::FindNormal(const std::string &n) const {
    auto it = normals.find(n);
    return it == normals.end() ? nullptr : &it->second;
}

// ---------------------------------------------------------------- tokenizer --
namespace {

struct Token {
    enum Type { END, IDENT, STRING, NUMBER, LBRACK, RBRACK } type;
    std::string text;
    float num;
};

struct Lexer {
    std::string buf;
    size_t pos = 0;
    std::string file;
    int line = 1;
    bool Load(const std::string &fn) {
        std::stringstream ss;
        if (fn == "-") ss << std::cin.rdbuf();   // ParseFile("-"): the scene on standard input (core/parser.cpp:58-60)
        else {
            std::ifstream f(fn, std::ios::binary);
            if (!f) return false;
            ss << f.rdbuf();
        }
        buf = ss.str();
        file = fn;
        return true;
    }
    Token Next() {
        Token t;
        t.num = 0;
        for (;;) {
            while (pos < buf.size() && isspace((unsigned char)buf[pos])) {
                if (buf[pos] == '\n') ++line;
                ++pos;
            }
            if (pos < buf.size() && buf[pos] == '#') {
                while (pos < buf.size() && buf[pos] != '\n') ++pos;
                continue;
            }
            break;
        }
        if (pos >= buf.size()) { t.type = Token::END; return t; }
        char c = buf[pos];
        if (c == '[') { ++pos; t.type = Token::LBRACK; return t; }
        if (c == ']') { ++pos; t.type = Token::RBRACK; return t; }
        if (c == '"') {
            size_t e = buf.find('"', pos + 1);
            if (e == std::string::npos) e = buf.size();
            t.type = Token::STRING;
            t.text = buf.substr(pos + 1, e - pos - 1);
            pos = e + 1;
            return t;
        }
        if (isalpha((unsigned char)c) || c == '_') {
            size_t s = pos;
            while (pos < buf.size() && (isalnum((unsigned char)buf[pos]) || buf[pos] == '_')) ++pos;
            t.type = Token::IDENT;
            t.text = buf.substr(s, pos - s);
            return t;
        }
        // NUMBER [-+]?([0-9]+|(([0-9]+\.[0-9]*)|(\.[0-9]+)))([eE][-+]?[0-9]+)?   pbrtlex.ll:114
        size_t s = pos;
        if (buf[pos] == '-' || buf[pos] == '+') ++pos;
        while (pos < buf.size() && (isdigit((unsigned char)buf[pos]) || buf[pos] == '.')) ++pos;
        if (pos < buf.size() && (buf[pos] == 'e' || buf[pos] == 'E')) {
            size_t save = pos;
            ++pos;
            if (pos < buf.size() && (buf[pos] == '-' || buf[pos] == '+')) ++pos;
            if (pos < buf.size() && isdigit((unsigned char)buf[pos])) {
                while (pos < buf.size() && isdigit((unsigned char)buf[pos])) ++pos;
            } else {
                pos = save;
            }
        }
        if (pos == s) {   // unknown character
            t.type = Token::IDENT;
            t.text = std::string(1, buf[pos++]);
            return t;
        }
        t.type = Token::NUMBER;
        t.text = buf.substr(s, pos - s);
        t.num = (float)atof(t.text.c_str());
        return t;
    }
};

// TransformCache (core/api.cpp:117-150): one Transform object per distinct CTM
struct TransformCache {
    std::map<Transform, std::pair<Transform *, Transform *>> cache;
    std::vector<std::unique_ptr<Transform>> store;
    void Lookup(const Transform &t, const Transform **tCached, const Transform **tInv) {
        auto it = cache.find(t);
        if (it == cache.end()) {
            store.emplace_back(new Transform(t));
            Transform *tr = store.back().get();
            store.emplace_back(new Transform(Inverse(t)));
            Transform *ti = store.back().get();
            it = cache.insert(std::make_pair(t, std::make_pair(tr, ti))).first;
        }
        if (tCached) *tCached = it->second.first;
        if (tInv) *tInv = it->second.second;
    }
};

struct TransformSet {   // core/api.cpp:77-103 (MAX_TRANSFORMS == 2)
    Transform t[2];
    bool IsAnimated() const { return t[0] != t[1]; }
};
TransformSet InverseSet(const TransformSet &ts) {
    TransformSet r;
    r.t[0] = Inverse(ts.t[0]);
    r.t[1] = Inverse(ts.t[1]);
    return r;
}

struct GraphicsState {   // core/api.cpp:213-232
    std::string material = "matte";
    ParamSet materialParams;
    std::string areaLight;
    ParamSet areaLightParams;
    bool reverseOrientation = false;
    // Texture (api.cpp:910-936) for the classes whose value does not depend on the surface point -- "constant", and "scale" /
    // "mix" of such -- kept as that value: what Texture::Evaluate returns everywhere (textures/{constant,scale,mix}.h)
    std::map<std::string, float> floatTextures;
    std::map<std::string, V3> spectrumTextures;
    std::map<std::string, int> namedMaterials;   // MakeNamedMaterial (api.cpp:947-961): name -> index into scene->materials
    std::string currentNamedMaterial;            // NamedMaterial (api.cpp:965-968); Material resets it (:939-944)
};

struct Api {
    // render options (core/api.cpp:153-209)
    float transformStartTime = 0.f, transformEndTime = 1.f;
    std::string FilterName = "box", FilmName = "image", SamplerName = "lowdiscrepancy",
                AcceleratorName = "bvh", RendererName = "sampler", SurfIntegratorName = "directlighting",
                VolIntegratorName = "emission", CameraName = "perspective";
    ParamSet FilterParams, FilmParams, SamplerParams, AcceleratorParams, RendererParams,
        SurfIntegratorParams, VolIntegratorParams, CameraParams;
    TransformSet CameraToWorld;
    // state
    TransformSet curTransform;
    int activeBits = 3;
    std::map<std::string, TransformSet> namedCoordinateSystems;
    GraphicsState graphicsState;
    std::vector<GraphicsState> pushedGraphicsStates;
    std::vector<TransformSet> pushedTransforms;
    std::vector<int> pushedActiveBits;
    bool inWorld = false, worldEnded = false;
    // ObjectBegin / ObjectEnd / ObjectInstance (core/api.cpp:1095-1147)
    struct InstanceDef {
        std::vector<GeometricPrimitive> prims;
        std::shared_ptr<GpuBVHAccel> object;   // built by the first ObjectInstance, shared by all of them
    };
    std::map<std::string, InstanceDef> instances;
    InstanceDef *currentInstance = nullptr;
    // scene under construction
    std::unique_ptr<Scene> scene;
    std::vector<GeometricPrimitive> primitives;
    std::vector<std::unique_ptr<Transform>> lightXforms;
    static TransformCache &Cache() { static TransformCache c; return c; }
    std::string baseDir;
    std::string error;
    Options opt;

    bool Fail(const std::string &msg) {
        if (error.empty()) error = msg;
        return false;
    }
    template <typename F> void ForActive(F f) {
        for (int i = 0; i < 2; ++i)
            if (activeBits & (1 << i)) f(curTransform.t[i]);
    }
    int MakeMaterial(const ParamSet &geomParams);
    int MakeMaterialOfType(const std::string &name, const ParamSet &geomParams, const ParamSet &mp);
    bool DoShape(const std::string &name, const ParamSet &params);
    bool DoLight(const std::string &name, const ParamSet &params);
    bool Finish(LoadedScene *out);
};

std::string ResolvePath(const std::string &baseDir, const std::string &fn) {   // core/fileutil.cpp
    if (fn.empty() || fn[0] == '/' || baseDir.empty()) return fn;
    return baseDir + "/" + fn;
}

// TextureParams lookup order: shape parameters first, then material parameters
// (core/paramset.cpp:600-640)
// TextureParams::GetSpectrumTexture / GetFloatTexture (core/paramset.cpp:598-630) evaluated: a named texture if the
// parameter is given as one ("texture Kd" "name": shape parameters first), else the constant; a name that was never
// defined is an Error message and then the constant, as in the reference
static thread_local const GraphicsState *g_texState = nullptr;   // the graphics state the lookups below read (set around MakeMaterialOfType / Texture)
static std::string FindTexture(const ParamSet &ps, const std::string &n) {
    auto it = ps.textures.find(n);
    return it == ps.textures.end() ? std::string() : it->second;
}
V3 TexSpectrum(const ParamSet &geom, const ParamSet &mat, const std::string &n, const V3 &def) {
    std::string name = FindTexture(geom, n);
    if (name.empty()) name = FindTexture(mat, n);
    if (!name.empty()) {
        if (g_texState && g_texState->spectrumTextures.count(name)) return g_texState->spectrumTextures.at(name);
        fprintf(stderr, "Error: Couldn't find spectrum texture named \"%s\" for parameter \"%s\"\n", name.c_str(), n.c_str());
    }
    return geom.FindOneSpectrum(n, mat.FindOneSpectrum(n, def));
}
float TexFloat(const ParamSet &geom, const ParamSet &mat, const std::string &n, float def) {
    std::string name = FindTexture(geom, n);
    if (name.empty()) name = FindTexture(mat, n);
    if (!name.empty()) {
        if (g_texState && g_texState->floatTextures.count(name)) return g_texState->floatTextures.at(name);
        fprintf(stderr, "Error: Couldn't find float texture named \"%s\" for parameter \"%s\"\n", name.c_str(), n.c_str());
    }
    return geom.FindOneFloat(n, mat.FindOneFloat(n, def));
}

int Api::MakeMaterial(const ParamSet &geomParams) {   // GraphicsState::CreateMaterial, core/api.cpp:1062-1077
    if (!graphicsState.currentNamedMaterial.empty()) {
        auto it = graphicsState.namedMaterials.find(graphicsState.currentNamedMaterial);
        if (it != graphicsState.namedMaterials.end()) return it->second;
    }
    return MakeMaterialOfType(graphicsState.material, geomParams, graphicsState.materialParams);
}

int Api::MakeMaterialOfType(const std::string &name, const ParamSet &geomParams, const ParamSet &mp) {   // MakeMaterial, api.cpp:317-372
    Material m;
    memset(&m.desc, 0, sizeof(m.desc));
    m.desc.measured = -1;
    g_texState = &graphicsState;
    if (geomParams.floats.count("bumpmap") || mp.floats.count("bumpmap") || geomParams.textures.count("bumpmap") || mp.textures.count("bumpmap")) {
        Fail("bump maps are out of scope");
        return -1;
    }
    if (name == "matte") {   // materials/matte.cpp:66-72
        m.desc.type = PBRTGPU_MAT_MATTE;
        V3 kd = TexSpectrum(geomParams, mp, "Kd", V3(0.5f, 0.5f, 0.5f));
        m.desc.Kd[0] = kd.x; m.desc.Kd[1] = kd.y; m.desc.Kd[2] = kd.z;
        m.desc.sigma = TexFloat(geomParams, mp, "sigma", 0.f);
        if (Clampf(m.desc.sigma, 0.f, 90.f) != 0.f) { Fail("matte sigma != 0 (Oren-Nayar) is outside the hot-path scope"); return -1; }
    } else if (name == "plastic") {   // materials/plastic.cpp:69-76
        m.desc.type = PBRTGPU_MAT_PLASTIC;
        V3 kd = TexSpectrum(geomParams, mp, "Kd", V3(0.25f, 0.25f, 0.25f));
        V3 ks = TexSpectrum(geomParams, mp, "Ks", V3(0.25f, 0.25f, 0.25f));
        m.desc.Kd[0] = kd.x; m.desc.Kd[1] = kd.y; m.desc.Kd[2] = kd.z;
        m.desc.Ks[0] = ks.x; m.desc.Ks[1] = ks.y; m.desc.Ks[2] = ks.z;
        m.desc.roughness = TexFloat(geomParams, mp, "roughness", .1f);
    } else if (name == "mirror") {   // materials/mirror.cpp:56-61
        m.desc.type = PBRTGPU_MAT_MIRROR;
        V3 kr = TexSpectrum(geomParams, mp, "Kr", V3(0.9f, 0.9f, 0.9f));
        m.desc.Ks[0] = kr.x; m.desc.Ks[1] = kr.y; m.desc.Ks[2] = kr.z;
    } else if (name == "glass") {    // materials/glass.cpp:60-67
        m.desc.type = PBRTGPU_MAT_GLASS;
        V3 kr = TexSpectrum(geomParams, mp, "Kr", V3(1.f, 1.f, 1.f));
        V3 kt = TexSpectrum(geomParams, mp, "Kt", V3(1.f, 1.f, 1.f));
        m.desc.Ks[0] = kr.x; m.desc.Ks[1] = kr.y; m.desc.Ks[2] = kr.z;
        m.desc.Kd[0] = kt.x; m.desc.Kd[1] = kt.y; m.desc.Kd[2] = kt.z;
        m.desc.roughness = TexFloat(geomParams, mp, "index", 1.5f);
    } else if (name == "metal") {    // materials/metal.cpp:101-112
        m.desc.type = PBRTGPU_MAT_METAL;
        // defaults: RGBSpectrum::FromSampled of copper's measured n and k (metal.cpp:70-105) -- the RGB coefficients the
        // reference computes, read out of it by the test tool `refdump metal` (spectrum-typed parameters, i.e. sampled
        // SPDs, are not read by this front end: "rgb" / "color" values or these defaults)
        V3 eta = TexSpectrum(geomParams, mp, "eta", V3(0x1.9a73d8p-3f, 0x1.d91c9cp-1f, 0x1.1a2ad8p+0f));
        V3 k = TexSpectrum(geomParams, mp, "k", V3(0x1.f4db58p+1f, 0x1.39f6acp+1f, 0x1.1233ap+1f));
        m.desc.Kd[0] = eta.x; m.desc.Kd[1] = eta.y; m.desc.Kd[2] = eta.z;
        m.desc.Ks[0] = k.x; m.desc.Ks[1] = k.y; m.desc.Ks[2] = k.z;
        m.desc.roughness = TexFloat(geomParams, mp, "roughness", .01f);
    } else if (name == "measured") {   // materials/measured.cpp:213-217
        m.desc.type = PBRTGPU_MAT_MEASURED;
        std::string fn = geomParams.FindOneString("filename", mp.FindOneString("filename", ""));
        fn = ResolvePath(baseDir, fn);
        size_t dot = fn.rfind('.');
        std::string suffix = dot == std::string::npos ? "" : fn.substr(dot);
        if (suffix != ".brdf" && suffix != ".BRDF") { Fail("measured material: only .brdf (irregular isotropic) data is in scope: " + fn); return -1; }
        int idx = -1;
        for (size_t i = 0; i < scene->measuredNames.size(); ++i)
            if (scene->measuredNames[i] == fn) idx = (int)i;
        if (idx < 0) {   // loadedThetaPhi cache, measured.cpp:96-99
            MeasuredData md;
            std::string err;
            if (!LoadIrregIsotropicBRDF(fn, &md, &err)) { Fail(err); return -1; }
            scene->measured.push_back(std::move(md));
            scene->measuredNames.push_back(fn);
            idx = (int)scene->measured.size() - 1;
        }
        m.desc.measured = idx;
    } else {
        Fail("material \"" + name + "\" is outside the hot-path scope (matte, plastic, measured, mirror, glass, metal)");
        return -1;
    }
    // identical descriptors share one slot
    for (size_t i = 0; i < scene->materials.size(); ++i)
        if (memcmp(&scene->materials[i].desc, &m.desc, sizeof(m.desc)) == 0) return (int)i;
    scene->materials.push_back(m);
    return (int)scene->materials.size() - 1;
}

// "string binfile": the mesh arrays of a trianglemesh / loopsubdiv from a binary file instead of the parameter list (SURVEY 8(f)
// rank 3: no text parsing, no per-number allocations for multi-million-triangle meshes).  Not a reference format: the arrays
// land in the same ParamSet slots the text form fills ("indices", "P", "N", "uv"), so everything downstream is the same code.
// Layout (little-endian): "PBRTMESH", uint32 flags (1: normals, 2: uvs), uint32 nIndices, uint32 nVertices, int32 indices[],
// float P[3 nv], float N[3 nv] if flagged, float uv[2 nv] if flagged.  pbrt_v2_b200.write_binary_mesh writes it.
static bool LoadBinaryMesh(const std::string &fn, ParamSet *ps, std::string *err) {
    FILE *f = fopen(fn.c_str(), "rb");
    if (!f) { *err = "cannot open binary mesh " + fn; return false; }
    char magic[8];
    uint32_t head[3];
    bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, "PBRTMESH", 8) == 0 && fread(head, 4, 3, f) == 3;
    const uint32_t flags = head[0], ni = head[1], nv = head[2];
    if (ok && (ni > (1u << 31) - 1 || nv > (1u << 30) || (flags & ~3u))) ok = false;
    if (ok) {   // the arrays must be there before memory is set aside for them
        const unsigned long long need = 4ull * ni + 12ull * nv + ((flags & 1u) ? 12ull * nv : 0) + ((flags & 2u) ? 8ull * nv : 0);
        const long at = ftell(f);
        fseek(f, 0, SEEK_END);
        const long end = ftell(f);
        fseek(f, at, SEEK_SET);
        if (at < 0 || end < at || (unsigned long long)(end - at) < need) ok = false;
    }
    if (ok) {
        std::vector<int> &idx = ps->ints["indices"];
        idx.resize(ni);
        ok = fread(idx.data(), 4, ni, f) == ni;
        auto readV3 = [&](std::vector<V3> *dst) {
            std::vector<float> raw((size_t)3 * nv);
            if (fread(raw.data(), 4, raw.size(), f) != raw.size()) return false;
            dst->resize(nv);
            for (uint32_t i = 0; i < nv; ++i) (*dst)[i] = V3(raw[3 * i], raw[3 * i + 1], raw[3 * i + 2]);
            return true;
        };
        ok = ok && readV3(&ps->points["P"]);
        if (ok && (flags & 1u)) ok = readV3(&ps->normals["N"]);
        if (ok && (flags & 2u)) {
            std::vector<float> &uv = ps->floats["uv"];
            uv.resize((size_t)2 * nv);
            ok = fread(uv.data(), 4, uv.size(), f) == uv.size();
        }
    }
    fclose(f);
    if (!ok) *err = "malformed binary mesh " + fn;
    return ok;
}

bool Api::DoShape(const std::string &name, const ParamSet &paramsIn) {   // core/api.cpp:990-1059
    ParamSet withMesh;
    const std::string binfile = paramsIn.FindOneString("binfile", "");
    if (!binfile.empty() && (name == "trianglemesh" || name == "loopsubdiv")) {
        withMesh = paramsIn;
        std::string err;
        if (!LoadBinaryMesh(ResolvePath(baseDir, binfile), &withMesh, &err)) return Fail(err);
    }
    const ParamSet &params = binfile.empty() ? paramsIn : withMesh;
    if (params.textures.count("alpha") || params.floats.count("alpha"))   // trianglemesh.cpp:397-406: alpha-tested triangles
        return Fail("alpha textures on shapes are outside the hot-path scope");
    if (name == "trianglemesh" && params.vectors.count("S"))   // trianglemesh.cpp:392-396: they turn the shading frame (:254-266)
        return Fail("per-vertex tangents (\"vector S\") on triangle meshes are outside the hot-path scope");
    const bool animated = curTransform.IsAnimated();
    const Transform *o2w, *w2o;
    if (!animated) Cache().Lookup(curTransform.t[0], &o2w, &w2o);
    else {   // api.cpp:1017-1020: the shape of an animated primitive is made with the identity
        Cache().Lookup(Transform(), &o2w, nullptr);
        w2o = o2w;
    }
    bool ro = graphicsState.reverseOrientation;
    std::shared_ptr<Shape> shape;
    if (name == "trianglemesh") {
        shape.reset(CreateTriangleMeshShape(o2w, w2o, ro, params));
    } else if (name == "heightfield") {
        // shapes/heightfield.cpp:120-130 + Heightfield::Refine (:69-117): nu x nv heights over the unit square become ONE
        // triangle mesh (two triangles per cell, uv = xy) -- which is all FullyRefine ever sees of the shape
        const int nu = params.FindOneInt("nu", -1), nv = params.FindOneInt("nv", -1);
        const std::vector<float> *Pz = params.FindFloat("Pz");
        if (nu < 2 || nv < 2 || !Pz || (long long)Pz->size() != (long long)nu * nv || (long long)nu * nv > (1ll << 26))
            return Fail("heightfield: \"integer nu\", \"integer nv\" (at least 2 each) and nu x nv \"float Pz\" values expected");
        ParamSet mesh;
        std::vector<V3> &P = mesh.points["P"];
        std::vector<float> &uv = mesh.floats["uv"];
        std::vector<int> &idx = mesh.ints["indices"];
        P.resize((size_t)nu * nv);
        uv.resize((size_t)2 * nu * nv);
        int pos = 0;
        for (int y = 0; y < nv; ++y)
            for (int x = 0; x < nu; ++x) {
                P[pos].x = uv[2 * pos] = (float)x / (float)(nu - 1);
                P[pos].y = uv[2 * pos + 1] = (float)y / (float)(nv - 1);
                P[pos].z = (*Pz)[pos];
                ++pos;
            }
        idx.reserve((size_t)6 * (nu - 1) * (nv - 1));
        auto vert = [nu](int x, int y) { return x + y * nu; };
        for (int y = 0; y < nv - 1; ++y)
            for (int x = 0; x < nu - 1; ++x) {
                idx.push_back(vert(x, y)); idx.push_back(vert(x + 1, y)); idx.push_back(vert(x + 1, y + 1));
                idx.push_back(vert(x, y)); idx.push_back(vert(x + 1, y + 1)); idx.push_back(vert(x, y + 1));
            }
        shape.reset(CreateTriangleMeshShape(o2w, w2o, ro, mesh));
    } else if (name == "sphere") {   // shapes/sphere.cpp:216-224
        float radius = params.FindOneFloat("radius", 1.f);
        float zmin = params.FindOneFloat("zmin", -radius);
        float zmax = params.FindOneFloat("zmax", radius);
        float phimax = params.FindOneFloat("phimax", 360.f);
        shape.reset(new Sphere(o2w, w2o, ro, radius, zmin, zmax, phimax));
    } else if (name == "disk") {   // shapes/disk.cpp:139-146
        float height = params.FindOneFloat("height", 0.);
        float radius = params.FindOneFloat("radius", 1);
        float inner = params.FindOneFloat("innerradius", 0);
        float phimax = params.FindOneFloat("phimax", 360);
        shape.reset(new Disk(o2w, w2o, ro, height, radius, inner, phimax));
    } else if (name == "loopsubdiv") {   // shapes/loopsubdiv.cpp:492-505
        int nlevels = params.FindOneInt("nlevels", 3);
        const std::vector<int> *vi = params.FindInt("indices");
        const std::vector<V3> *P = params.FindPoint("P");
        if (vi && P) {
            for (int idx : *vi)   // (the reference does not check, loopsubdiv.cpp:166-174, and reads out of bounds)
                if (idx < 0 || idx >= (int)P->size()) return Fail("loopsubdiv has an out-of-bounds vertex index");
            LoopSubdiv *ls = new LoopSubdiv(o2w, w2o, ro, (int)vi->size() / 3, (int)P->size(), vi->data(), P->data(), nlevels);
            // "string refiner" [ "host" | "gpu" ] (also PBRTGPU_SUBDIV in the environment): where Refine() runs
            const char *env = getenv("PBRTGPU_SUBDIV");
            const std::string refiner = params.FindOneString("refiner", env ? env : "host");
            if (refiner != "host" && refiner != "gpu")
                fprintf(stderr, "Warning: loopsubdiv refiner \"%s\" unknown.  Using \"host\".\n", refiner.c_str());
            ls->refineOnDevice = refiner == "gpu";
            shape.reset(ls);
        }
    } else {
        return Fail("shape \"" + name + "\" is outside the hot-path scope");
    }
    if (!shape) return true;   // MakeShape returned NULL: ignored like api.cpp:1001
    int mtl = MakeMaterial(params);
    if (mtl < 0) return false;
    if (currentInstance && animated)
        return Fail("animated shapes inside an object definition (a TransformedPrimitive inside a TransformedPrimitive) are outside the hot-path scope");
    if (currentInstance && !graphicsState.areaLight.empty())
        // (the reference warns "Area lights not supported with object instancing" and keeps a light that emits when hit but
        // is never sampled, api.cpp:1046-1050: there is no such light in the descriptor)
        return Fail("area lights inside an object definition are outside the hot-path scope");
    if (animated) {   // api.cpp:1011-1043
        if (!graphicsState.areaLight.empty())
            fprintf(stderr, "Warning: Ignoring currently set area light when creating animated shape\n");
        const Transform *w2o0, *w2o1;
        Cache().Lookup(curTransform.t[0], nullptr, &w2o0);
        Cache().Lookup(curTransform.t[1], nullptr, &w2o1);
        AnimatedTransform animatedWorldToObject(*w2o0, transformStartTime, *w2o1, transformEndTime);
        GeometricPrimitive base;
        base.shape = shape;
        base.material = mtl;
        base.areaLight = -1;
        GeometricPrimitive gp;
        gp.animated.reset(new TransformedPrimitive(base, animatedWorldToObject));
        if (!gp.animated->primitive->buildError.empty()) return Fail(gp.animated->primitive->buildError);
        if (gp.animated->primitive->primitives.empty()) return true;   // api.cpp:1037
        primitives.push_back(gp);
        return true;
    }
    int area = -1;
    if (!graphicsState.areaLight.empty()) {   // api.cpp:1005-1008, lights/diffuse.cpp:58-66
        if (graphicsState.areaLight != "area" && graphicsState.areaLight != "diffuse")
            return Fail("area light \"" + graphicsState.areaLight + "\" unknown");
        if (shape->kind != Shape::SPHERE && shape->kind != Shape::DISK)
            return Fail("area lights on shapes other than sphere/disk are outside the hot-path scope");
        const ParamSet &ap = graphicsState.areaLightParams;
        V3 L = ap.FindOneSpectrum("L", V3(1, 1, 1));
        V3 sc = ap.FindOneSpectrum("scale", V3(1, 1, 1));
        Light lt;
        lt.desc.type = PBRTGPU_LIGHT_DIFFUSE_AREA;
        int nSamples = ap.FindOneInt("nsamples", 1);
        if (opt.quickRender) nSamples = std::max(1, nSamples / 4);   // diffuse.cpp:64
        lt.desc.nSamples = std::max(1, nSamples);   // diffuse.cpp:63, light.h:53
        lt.desc.L[0] = L.x * sc.x; lt.desc.L[1] = L.y * sc.y; lt.desc.L[2] = L.z * sc.z;
        lt.shape = shape.get();
        scene->lights.push_back(lt);
        area = (int)scene->lights.size() - 1;
    }
    GeometricPrimitive gp;
    gp.shape = shape;
    gp.material = mtl;
    gp.areaLight = area;
    if (currentInstance) currentInstance->prims.push_back(gp);   // api.cpp:1046-1051
    else primitives.push_back(gp);
    return true;
}

bool Api::DoLight(const std::string &name, const ParamSet &params) {   // core/api.cpp:956-967
    if (name == "infinite" || name == "exinfinite") {   // lights/infinite.cpp:256-265 (api.cpp:549-551: both names)
        V3 L = params.FindOneSpectrum("L", V3(1, 1, 1));
        V3 sc = params.FindOneSpectrum("scale", V3(1, 1, 1));
        Light lt;
        lt.desc.type = PBRTGPU_LIGHT_INFINITE;
        int nSamples = params.FindOneInt("nsamples", 1);
        if (opt.quickRender) nSamples = std::max(1, nSamples / 4);   // infinite.cpp:190
        lt.desc.nSamples = std::max(1, nSamples);   // Light ctor, light.cpp:46-51
        lt.desc.L[0] = L.x * sc.x; lt.desc.L[1] = L.y * sc.y; lt.desc.L[2] = L.z * sc.z;
        // "mapname": ReadImage (core/imageio.cpp:58-81): .pfm, .tga and .exr (envmap.cpp, tga.cpp, exr.cpp)
        std::string mapname = params.FindOneString("mapname", "");
        std::vector<float> rgb;
        int w = 1, h = 1;
        if (!mapname.empty()) {
            mapname = ResolvePath(baseDir, mapname);
            std::string err;
            if (!ReadImage(mapname, &w, &h, &rgb, &err)) return Fail("infinite light: " + err);
        }
        lt.envMap = std::make_shared<EnvMap>();
        BuildEnvMap(lt.desc.L, w, h, mapname.empty() ? nullptr : rgb.data(), lt.envMap.get());
        lt.lightToWorld = curTransform.t[0];
        scene->lights.push_back(lt);
        return true;
    }
    if (name != "point") return Fail("light \"" + name + "\" is outside the hot-path scope (point, area, infinite)");
    // lights/point.cpp:68-75 and :41-46
    V3 I = params.FindOneSpectrum("I", V3(1, 1, 1));
    V3 sc = params.FindOneSpectrum("scale", V3(1, 1, 1));
    V3 P = params.FindOnePoint("from", V3(0, 0, 0));
    Transform l2w = Translate(V3(P.x, P.y, P.z)) * curTransform.t[0];
    Light lt;
    lt.desc.type = PBRTGPU_LIGHT_POINT;
    lt.desc.nSamples = 1;   // Light(l2w) default, light.h:52
    lt.desc.L[0] = I.x * sc.x; lt.desc.L[1] = I.y * sc.y; lt.desc.L[2] = I.z * sc.z;
    V3 pos = l2w.Pt(V3(0, 0, 0));
    lt.desc.pos[0] = pos.x; lt.desc.pos[1] = pos.y; lt.desc.pos[2] = pos.z;
    scene->lights.push_back(lt);
    return true;
}

bool Api::Finish(LoadedScene *out) {   // core/api.cpp:1150-1281
    // --- film + filter + camera (api.cpp:1272-1281, film/image.cpp:220-258) ---
    // MakeFilter, core/api.cpp:572-589; filters/{box,gaussian,mitchell,sinc,triangle}.cpp.  Filters wider than half a pixel
    // reach the kernels as ImageFilm's weight table (PbrtGpuRenderDesc::filterTable; the shipped build refuses them)
    float fxw, fyw;
    std::function<float(float, float)> evaluate;
    if (FilterName == "box") {
        fxw = FilterParams.FindOneFloat("xwidth", 0.5f); fyw = FilterParams.FindOneFloat("ywidth", 0.5f);
        evaluate = [](float, float) { return 1.f; };
    } else if (FilterName == "gaussian") {   // gaussian.h:47-58
        fxw = FilterParams.FindOneFloat("xwidth", 2.f); fyw = FilterParams.FindOneFloat("ywidth", 2.f);
        const float alpha = FilterParams.FindOneFloat("alpha", 2.f);
        const float expX = expf(-alpha * fxw * fxw), expY = expf(-alpha * fyw * fyw);
        auto gaussian = [alpha](float d, float expv) { return std::max(0.f, float(expf(-alpha * d * d) - expv)); };
        evaluate = [=](float x, float y) { return gaussian(x, expX) * gaussian(y, expY); };
    } else if (FilterName == "mitchell") {   // mitchell.h:47-63
        fxw = FilterParams.FindOneFloat("xwidth", 2.f); fyw = FilterParams.FindOneFloat("ywidth", 2.f);
        const float B = FilterParams.FindOneFloat("B", 1.f / 3.f), C = FilterParams.FindOneFloat("C", 1.f / 3.f);
        const float invX = 1.f / fxw, invY = 1.f / fyw;
        auto mitchell1D = [B, C](float x) {
            x = fabsf(2.f * x);
            if (x > 1.f)
                return ((-B - 6 * C) * x * x * x + (6 * B + 30 * C) * x * x + (-12 * B - 48 * C) * x + (8 * B + 24 * C)) * (1.f / 6.f);
            else
                return ((12 - 9 * B - 6 * C) * x * x * x + (-18 + 12 * B + 6 * C) * x * x + (6 - 2 * B)) * (1.f / 6.f);
        };
        evaluate = [=](float x, float y) { return mitchell1D(x * invX) * mitchell1D(y * invY); };
    } else if (FilterName == "sinc") {       // sinc.h:47-59
        fxw = FilterParams.FindOneFloat("xwidth", 4.f); fyw = FilterParams.FindOneFloat("ywidth", 4.f);
        const float tau = FilterParams.FindOneFloat("tau", 3.f);
        const float invX = 1.f / fxw, invY = 1.f / fyw;
        auto sinc1D = [tau](float x) {
            x = fabsf(x);
            if (x < 1e-5) return 1.f;
            if (x > 1.) return 0.f;
            x *= kPi;
            float sinc = sinf(x) / x;
            float lanczos = sinf(x * tau) / (x * tau);
            return sinc * lanczos;
        };
        evaluate = [=](float x, float y) { return sinc1D(x * invX) * sinc1D(y * invY); };
    } else if (FilterName == "triangle") {   // triangle.cpp:39-42
        fxw = FilterParams.FindOneFloat("xwidth", 2.f); fyw = FilterParams.FindOneFloat("ywidth", 2.f);
        const float xw = fxw, yw = fyw;
        evaluate = [=](float x, float y) { return std::max(0.f, xw - fabsf(x)) * std::max(0.f, yw - fabsf(y)); };
    } else
        return Fail("pixel filter \"" + FilterName + "\" unknown");
    if (FilmName != "image") return Fail("film \"" + FilmName + "\" unknown");
    std::string filename = FilmParams.FindOneString("filename", "");
    if (!opt.imageFile.empty() && filename.empty()) filename = opt.imageFile;
    if (filename.empty()) filename = "pbrt.tga";   // film/image.cpp:267-271 (the default of a build without OpenEXR)
    int xres = FilmParams.FindOneInt("xresolution", 640);
    int yres = FilmParams.FindOneInt("yresolution", 480);
    if (opt.quickRender) xres = std::max(1, xres / 4);   // film/image.cpp:244-245
    if (opt.quickRender) yres = std::max(1, yres / 4);
    // (the reference takes any integers here and fails later, in int arithmetic or in an allocation)
    if (xres <= 0 || yres <= 0 || (long long)xres * yres > (1ll << 28)) return Fail("film resolution out of range");
    float crop[4] = {0, 1, 0, 1};
    const std::vector<float> *cr = FilmParams.FindFloat("cropwindow");
    if (cr && cr->size() == 4) {
        crop[0] = Clampf(std::min((*cr)[0], (*cr)[1]), 0., 1.);
        crop[1] = Clampf(std::max((*cr)[0], (*cr)[1]), 0., 1.);
        crop[2] = Clampf(std::min((*cr)[2], (*cr)[3]), 0., 1.);
        crop[3] = Clampf(std::max((*cr)[2], (*cr)[3]), 0., 1.);
    }
    std::unique_ptr<GpuSamplerRenderer> r(new GpuSamplerRenderer);
    r->nCores = opt.nCores;
    r->film.reset(new ImageFilm(xres, yres, fxw, fyw, crop, filename));
    for (int y = 0; y < 16; ++y) {   // film/image.cpp:57-69 (FILTER_TABLE_SIZE 16)
        float fy = ((float)y + .5f) * fyw / 16;
        for (int x = 0; x < 16; ++x) {
            float fx = ((float)x + .5f) * fxw / 16;
            r->film->filterTable[16 * y + x] = evaluate(fx, fy);
        }
    }
    if (CameraName != "perspective") return Fail("camera \"" + CameraName + "\" is out of scope (perspective)");
    if (CameraToWorld.IsAnimated()) return Fail("animated cameras are out of scope");
    {   // cameras/perspective.cpp:141-180
        float shutteropen = CameraParams.FindOneFloat("shutteropen", 0.f);
        float shutterclose = CameraParams.FindOneFloat("shutterclose", 1.f);
        if (shutterclose < shutteropen) std::swap(shutterclose, shutteropen);
        float lensradius = CameraParams.FindOneFloat("lensradius", 0.f);
        float focaldistance = CameraParams.FindOneFloat("focaldistance", 1e30f);
        float frame = CameraParams.FindOneFloat("frameaspectratio", float(xres) / float(yres));
        float screen[4];
        if (frame > 1.f) { screen[0] = -frame; screen[1] = frame; screen[2] = -1.f; screen[3] = 1.f; }
        else { screen[0] = -1.f; screen[1] = 1.f; screen[2] = -1.f / frame; screen[3] = 1.f / frame; }
        const std::vector<float> *sw = CameraParams.FindFloat("screenwindow");
        if (sw && sw->size() == 4) for (int i = 0; i < 4; ++i) screen[i] = (*sw)[i];
        float fov = CameraParams.FindOneFloat("fov", 90.);
        float halffov = CameraParams.FindOneFloat("halffov", -1.f);
        if (halffov > 0.f) fov = 2.f * halffov;
        r->camera.reset(new PerspectiveCamera(CameraToWorld.t[0], screen, shutteropen, shutterclose,
                                              lensradius, focaldistance, fov, r->film.get()));
    }
    // --- sampler (samplers/lowdiscrepancy.cpp:40-51,82-92) ---
    if (SamplerName != "lowdiscrepancy") return Fail("sampler \"" + SamplerName + "\" is out of scope (lowdiscrepancy)");
    {
        int nsamp = SamplerParams.FindOneInt("pixelsamples", 4);
        if (opt.quickRender) nsamp = 1;   // lowdiscrepancy.cpp:88
        uint32_t v = (uint32_t)nsamp;   // RoundUpPow2, core/pbrt.h:271-277
        v--; v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16; v++;
        r->sampler.samplesPerPixel = (int)v;
        r->film->GetSampleExtent(&r->sampler.xPixelStart, &r->sampler.xPixelEnd, &r->sampler.yPixelStart,
                                 &r->sampler.yPixelEnd);
    }
    // --- integrators (api.cpp:533-563) ---
    if (SurfIntegratorName == "path" || SurfIntegratorName == "gpupath")
        r->surfaceIntegrator.reset(new GpuPathIntegrator(SurfIntegratorParams.FindOneInt("maxdepth", 5)));
    else if (SurfIntegratorName == "directlighting") {   // directlighting.cpp:120-133
        int maxDepth = SurfIntegratorParams.FindOneInt("maxdepth", 5);
        std::string st = SurfIntegratorParams.FindOneString("strategy", "all");
        int strategy = PBRTGPU_STRATEGY_ALL;
        if (st == "one") strategy = PBRTGPU_STRATEGY_ONE;
        else if (st != "all")
            fprintf(stderr, "Warning: Strategy \"%s\" for direct lighting unknown. Using \"all\".\n", st.c_str());
        r->surfaceIntegrator.reset(new GpuDirectLightingIntegrator(strategy, maxDepth));
    } else
        return Fail("surface integrator \"" + SurfIntegratorName + "\" is outside the scope (path, directlighting)");
    if (VolIntegratorName != "emission") return Fail("volume integrator \"" + VolIntegratorName + "\" is out of scope");
    // Sample::Sample (core/sampler.cpp:84-92): surface integrator's requests, then the volume integrator's
    r->surfaceIntegrator->RequestSamples(*scene, &r->n1D, &r->n2D);
    r->n1D.push_back(1);   // EmissionIntegrator::RequestSamples (emission.cpp:40-43): tauSampleOffset
    r->n1D.push_back(1);   //                                                       scatterSampleOffset
    if (r->n1D.size() > PBRTGPU_MAX_SAMPLE_ARRAYS || r->n2D.size() > PBRTGPU_MAX_SAMPLE_ARRAYS)
        return Fail("too many lights for the sample-array table of the C ABI (PBRTGPU_MAX_SAMPLE_ARRAYS)");
    if (RendererName != "sampler" && RendererName != "gpusampler")
        return Fail("renderer \"" + RendererName + "\" is out of scope (sampler)");
    if (RendererParams.FindOneBool("visualizeobjectids", false))   // api.cpp:1250, samplerrenderer.cpp:96-110: colours by primitive id
        return Fail("Renderer \"sampler\" \"bool visualizeobjectids\" is out of scope");
    // --- accelerator (api.cpp:580-594, 1186-1207) ---
    if (AcceleratorName != "bvh" && AcceleratorName != "gpubvh")
        return Fail("accelerator \"" + AcceleratorName + "\" is out of scope (bvh)");
    GpuBVHAccel *accel = CreateGpuBVHAccelerator(primitives, AcceleratorParams);
    scene->aggregate.reset(accel);
    if (!accel->buildError.empty()) return Fail(accel->buildError);
    std::string err;
    if (!scene->Flatten(&err)) return Fail(err);
    out->scene = std::move(scene);
    out->renderer = std::move(r);
    return true;
}

static thread_local std::string g_paramBaseDir;   // where "spectrum" parameters look for their .spd files: the scene file's directory

bool ParseParams(Lexer &lx, Token &tok, ParamSet *ps, std::string *err) {
    // after the directive's fixed arguments: ("type name" value | [values])*
    while (tok.type == Token::STRING) {
        std::istringstream decl(tok.text);
        std::string type, name;
        decl >> type >> name;
        if (name.empty()) { *err = "malformed parameter declaration \"" + tok.text + "\""; return false; }
        std::vector<float> nums;
        std::vector<std::string> strs;
        tok = lx.Next();
        auto take = [&](const Token &t) -> bool {
            if (t.type == Token::NUMBER) { nums.push_back(t.num); return true; }
            if (t.type == Token::STRING) { strs.push_back(t.text); return true; }
            return false;
        };
        if (tok.type == Token::LBRACK) {
            tok = lx.Next();
            while (tok.type != Token::RBRACK) {
                if (!take(tok)) { *err = "bad token in parameter array for \"" + name + "\""; return false; }
                tok = lx.Next();
            }
            tok = lx.Next();
        } else {
            if (!take(tok)) { *err = "missing value for parameter \"" + name + "\""; return false; }
            tok = lx.Next();
        }
        auto triples = [&](std::vector<V3> &dst) {
            for (size_t i = 0; i + 2 < nums.size(); i += 3) dst.push_back(V3(nums[i], nums[i + 1], nums[i + 2]));
        };
        if (type == "float") ps->floats[name] = nums;
        else if (type == "integer") {   // core/pbrtparse.yy: int(value)
            std::vector<int> v;
            for (float f : nums) v.push_back((int)f);
            ps->ints[name] = v;
        } else if (type == "bool") {
            std::vector<bool> v;
            for (const auto &s : strs) v.push_back(s == "true");
            ps->bools[name] = v;
        } else if (type == "point") triples(ps->points[name]);
        else if (type == "vector") triples(ps->vectors[name]);
        else if (type == "normal") triples(ps->normals[name]);
        else if (type == "color" || type == "rgb") triples(ps->spectra[name]);   // RGBSpectrum::FromRGB
        else if (type == "xyz") {   // ParamSet::AddXYZSpectrum (paramset.cpp:109-118): RGBSpectrum::FromXYZ = XYZToRGB
            std::vector<V3> xyz;
            triples(xyz);
            std::vector<V3> &dst = ps->spectra[name];
            dst.clear();
            for (const V3 &c : xyz)
                dst.push_back(V3(3.240479f * c.x - 1.537150f * c.y - 0.498535f * c.z, -0.969256f * c.x + 1.875991f * c.y + 0.041556f * c.z,
                                 0.055648f * c.x - 0.204043f * c.y + 1.057311f * c.z));
        } else if (type == "blackbody") {
            // ParamSet::AddBlackbodySpectrum (paramset.cpp:121-135): (temperature in K, scale) pairs; Blackbody() over the CIE
            // wavelengths (spectrum.cpp:195-204: Planck's law normalised at 555 nm, in double) -> FromSampled -> times scale
            if (nums.size() % 2) { *err = "blackbody \"" + name + "\": temperature / scale pairs expected"; return false; }
            std::vector<V3> &dst = ps->spectra[name];
            dst.clear();
            for (size_t j = 0; j + 1 < nums.size(); j += 2) {
                const float temp = nums[j], scale = nums[j + 1];
                float wl[471], v[471];
                for (int i = 0; i < 471; ++i) wl[i] = float(360 + i);   // CIE_lambda
                if (temp <= 0) {
                    for (int i = 0; i < 471; ++i) v[i] = 0.f;
                } else {
                    const double C2 = 1.4388e7;
                    double norm = pow(555.0, 5.0) * (exp(C2 / (555.0 * temp)) - 1.);
                    for (int i = 0; i < 471; ++i) v[i] = float(norm / (pow(double(wl[i]), 5.0) * (exp(C2 / (wl[i] * temp)) - 1.)));
                }
                float rgb[3];
                SortedSamplesToRGB(wl, v, 471, rgb);
                dst.push_back(V3(rgb[0] * scale, rgb[1] * scale, rgb[2] * scale));
            }
        } else if (type == "spectrum") {
            // ParamSet::AddSampledSpectrum / AddSampledSpectrumFiles (paramset.cpp:138-187): (wavelength, value) pairs, inline or
            // from .spd files next to the scene, through RGBSpectrum::FromSampled (spectrum.h:428-451; samples sorted first)
            auto fromPairs = [](const std::vector<float> &vals) {
                std::vector<std::pair<float, float>> pairs;
                for (size_t j = 0; j + 1 < vals.size(); j += 2) pairs.push_back(std::make_pair(vals[j], vals[j + 1]));
                bool sorted = true;
                for (size_t j = 0; j + 1 < pairs.size(); ++j) sorted = sorted && !(pairs[j].first > pairs[j + 1].first);
                if (!sorted) std::sort(pairs.begin(), pairs.end());   // SortSpectrumSamples, spectrum.cpp:45-55
                std::vector<float> wl, v;
                for (const auto &pr : pairs) { wl.push_back(pr.first); v.push_back(pr.second); }
                float rgb[3] = {0.f, 0.f, 0.f};
                if (!wl.empty()) SortedSamplesToRGB(wl.data(), v.data(), (int)wl.size(), rgb);
                return V3(rgb[0], rgb[1], rgb[2]);
            };
            std::vector<V3> &dst = ps->spectra[name];
            dst.clear();
            if (!strs.empty()) {
                for (const std::string &fnIn : strs) {
                    const std::string fn = ResolvePath(g_paramBaseDir, fnIn);
                    std::vector<float> vals;
                    if (!ReadSpectrumFile(fn, &vals)) {
                        fprintf(stderr, "Warning: Unable to read SPD file \"%s\".  Using black distribution.\n", fn.c_str());
                        dst.push_back(V3(0, 0, 0));
                    } else {
                        if (vals.size() % 2) fprintf(stderr, "Warning: Extra value found in spectrum file \"%s\". Ignoring it.\n", fn.c_str());
                        dst.push_back(fromPairs(vals));
                    }
                }
            } else {
                if (nums.size() % 2) { *err = "spectrum \"" + name + "\": wavelength / value pairs expected"; return false; }
                dst.push_back(fromPairs(nums));
            }
        }
        else if (type == "string") ps->strings[name] = strs;
        else if (type == "texture") ps->textures[name] = strs.empty() ? "" : strs[0];
        else { *err = "parameter type \"" + type + "\" is not supported by this front end"; return false; }
    }
    return true;
}

bool ParseStream(Api &api, const std::string &filename, int depth) {
    Lexer lx;
    if (!lx.Load(filename)) return api.Fail("unable to open scene file \"" + filename + "\"");
    if (depth > 32) return api.Fail("Include nesting too deep");
    Token tok = lx.Next();
    auto nums = [&](int n, float *out) -> bool {
        for (int i = 0; i < n; ++i) {
            if (tok.type != Token::NUMBER) return false;
            out[i] = tok.num;
            tok = lx.Next();
        }
        return true;
    };
    while (tok.type != Token::END) {
        if (tok.type != Token::IDENT) return api.Fail(filename + ": expected directive near line " + std::to_string(lx.line));
        std::string d = tok.text;
        tok = lx.Next();
        float v[16];
        std::string perr;
        auto named = [&](std::string *name, ParamSet *ps) -> bool {
            if (tok.type != Token::STRING) return api.Fail(d + ": expected a quoted name");
            *name = tok.text;
            tok = lx.Next();
            if (!ParseParams(lx, tok, ps, &perr)) return api.Fail(d + ": " + perr);
            return true;
        };
        if (d == "Identity") api.ForActive([&](Transform &t) { t = Transform(); });
        else if (d == "Translate") {
            if (!nums(3, v)) return api.Fail("Translate: expected 3 numbers");
            api.ForActive([&](Transform &t) { t = t * Translate(V3(v[0], v[1], v[2])); });
        } else if (d == "Scale") {
            if (!nums(3, v)) return api.Fail("Scale: expected 3 numbers");
            api.ForActive([&](Transform &t) { t = t * Scale(v[0], v[1], v[2]); });
        } else if (d == "Rotate") {
            if (!nums(4, v)) return api.Fail("Rotate: expected 4 numbers");
            api.ForActive([&](Transform &t) { t = t * Rotate(v[0], V3(v[1], v[2], v[3])); });
        } else if (d == "LookAt") {
            if (!nums(9, v)) return api.Fail("LookAt: expected 9 numbers");
            api.ForActive([&](Transform &t) { t = t * LookAt(V3(v[0], v[1], v[2]), V3(v[3], v[4], v[5]), V3(v[6], v[7], v[8]), nullptr); });
        } else if (d == "Transform" || d == "ConcatTransform") {
            if (tok.type != Token::LBRACK) return api.Fail(d + ": expected [");
            tok = lx.Next();
            if (!nums(16, v)) return api.Fail(d + ": expected 16 numbers");
            if (tok.type != Token::RBRACK) return api.Fail(d + ": expected ]");
            tok = lx.Next();
            Matrix4x4 m(v[0], v[4], v[8], v[12], v[1], v[5], v[9], v[13], v[2], v[6], v[10], v[14], v[3], v[7], v[11], v[15]);
            if (d == "Transform") api.ForActive([&](Transform &t) { t = Transform(m); });
            else api.ForActive([&](Transform &t) { t = t * Transform(m); });
        } else if (d == "CoordinateSystem") {
            if (tok.type != Token::STRING) return api.Fail("CoordinateSystem: expected name");
            api.namedCoordinateSystems[tok.text] = api.curTransform;
            tok = lx.Next();
        } else if (d == "CoordSysTransform") {
            if (tok.type != Token::STRING) return api.Fail("CoordSysTransform: expected name");
            auto it = api.namedCoordinateSystems.find(tok.text);
            if (it != api.namedCoordinateSystems.end()) api.curTransform = it->second;
            tok = lx.Next();
        } else if (d == "ActiveTransform") {
            if (tok.type != Token::IDENT) return api.Fail("ActiveTransform: expected All/StartTime/EndTime");
            if (tok.text == "All") api.activeBits = 3;
            else if (tok.text == "StartTime") api.activeBits = 1;
            else if (tok.text == "EndTime") api.activeBits = 2;
            else return api.Fail("ActiveTransform: bad argument");
            tok = lx.Next();
        } else if (d == "TransformTimes") {
            if (!nums(2, v)) return api.Fail("TransformTimes: expected 2 numbers");
            api.transformStartTime = v[0];
            api.transformEndTime = v[1];
        } else if (d == "Camera") {
            if (!named(&api.CameraName, &api.CameraParams)) return false;
            api.CameraToWorld = InverseSet(api.curTransform);
            api.namedCoordinateSystems["camera"] = api.CameraToWorld;
        } else if (d == "Film") { api.FilmParams = ParamSet(); if (!named(&api.FilmName, &api.FilmParams)) return false; }
        else if (d == "Sampler") { api.SamplerParams = ParamSet(); if (!named(&api.SamplerName, &api.SamplerParams)) return false; }
        else if (d == "PixelFilter") { api.FilterParams = ParamSet(); if (!named(&api.FilterName, &api.FilterParams)) return false; }
        else if (d == "Accelerator") { api.AcceleratorParams = ParamSet(); if (!named(&api.AcceleratorName, &api.AcceleratorParams)) return false; }
        else if (d == "SurfaceIntegrator") { api.SurfIntegratorParams = ParamSet(); if (!named(&api.SurfIntegratorName, &api.SurfIntegratorParams)) return false; }
        else if (d == "VolumeIntegrator") { api.VolIntegratorParams = ParamSet(); if (!named(&api.VolIntegratorName, &api.VolIntegratorParams)) return false; }
        else if (d == "Renderer") { api.RendererParams = ParamSet(); if (!named(&api.RendererName, &api.RendererParams)) return false; }
        else if (d == "WorldBegin") {
            api.inWorld = true;
            api.curTransform = TransformSet();
            api.activeBits = 3;
            api.namedCoordinateSystems["world"] = api.curTransform;
        } else if (d == "AttributeBegin") {
            api.pushedGraphicsStates.push_back(api.graphicsState);
            api.pushedTransforms.push_back(api.curTransform);
            api.pushedActiveBits.push_back(api.activeBits);
        } else if (d == "AttributeEnd") {
            if (api.pushedGraphicsStates.empty()) return api.Fail("unmatched AttributeEnd");
            api.graphicsState = api.pushedGraphicsStates.back(); api.pushedGraphicsStates.pop_back();
            api.curTransform = api.pushedTransforms.back(); api.pushedTransforms.pop_back();
            api.activeBits = api.pushedActiveBits.back(); api.pushedActiveBits.pop_back();
        } else if (d == "ObjectBegin") {   // api.cpp:1095-1102
            if (tok.type != Token::STRING) return api.Fail("ObjectBegin: expected a quoted name");
            api.pushedGraphicsStates.push_back(api.graphicsState);
            api.pushedTransforms.push_back(api.curTransform);
            api.pushedActiveBits.push_back(api.activeBits);
            if (api.currentInstance) return api.Fail("ObjectBegin called inside of instance definition");
            api.instances[tok.text] = Api::InstanceDef();
            api.currentInstance = &api.instances[tok.text];
            tok = lx.Next();
        } else if (d == "ObjectEnd") {   // api.cpp:1105-1111
            if (!api.currentInstance) return api.Fail("ObjectEnd called outside of instance definition");
            api.currentInstance = nullptr;
            if (api.pushedGraphicsStates.empty()) return api.Fail("unmatched ObjectEnd");
            api.graphicsState = api.pushedGraphicsStates.back(); api.pushedGraphicsStates.pop_back();
            api.curTransform = api.pushedTransforms.back(); api.pushedTransforms.pop_back();
            api.activeBits = api.pushedActiveBits.back(); api.pushedActiveBits.pop_back();
        } else if (d == "ObjectInstance") {   // api.cpp:1114-1147
            if (tok.type != Token::STRING) return api.Fail("ObjectInstance: expected a quoted name");
            const std::string name = tok.text;
            tok = lx.Next();
            if (api.currentInstance) return api.Fail("ObjectInstance can't be called inside instance definition");
            auto it = api.instances.find(name);
            if (it == api.instances.end()) return api.Fail("Unable to find instance named \"" + name + "\"");
            Api::InstanceDef &in = it->second;
            if (in.prims.empty() && !in.object) continue;
            if (!in.object) {
                // "Refine instance Primitives and create aggregate" unless it is one intersectable primitive (a sphere or a
                // disk here): the scene's Accelerator and its parameters (api.cpp:1127-1136)
                const bool single = in.prims.size() == 1 && in.prims[0].shape->kind != Shape::TRIANGLE_MESH &&
                                    in.prims[0].shape->kind != Shape::LOOP_SUBDIV;
                in.object.reset(CreateGpuBVHAccelerator(in.prims, api.AcceleratorParams));
                if (!in.object->buildError.empty()) return api.Fail(in.object->buildError);
                if (single) in.object->nodes.clear();
                for (const GeometricPrimitive &g : in.prims) in.object->owned.push_back(g.shape);   // the refined primitives point into them
                in.prims.clear();
                if (in.object->primitives.empty()) { in.object.reset(); continue; }
            }
            const Transform *w2i0, *w2i1;
            api.Cache().Lookup(api.curTransform.t[0], nullptr, &w2i0);
            api.Cache().Lookup(api.curTransform.t[1], nullptr, &w2i1);
            AnimatedTransform animatedWorldToInstance(*w2i0, api.transformStartTime, *w2i1, api.transformEndTime);
            GeometricPrimitive gp;
            gp.animated.reset(new TransformedPrimitive(in.object, animatedWorldToInstance));
            api.primitives.push_back(gp);
        } else if (d == "TransformBegin") {
            api.pushedTransforms.push_back(api.curTransform);
            api.pushedActiveBits.push_back(api.activeBits);
        } else if (d == "TransformEnd") {
            if (api.pushedTransforms.empty()) return api.Fail("unmatched TransformEnd");
            api.curTransform = api.pushedTransforms.back(); api.pushedTransforms.pop_back();
            api.activeBits = api.pushedActiveBits.back(); api.pushedActiveBits.pop_back();
        } else if (d == "ReverseOrientation") {
            api.graphicsState.reverseOrientation = !api.graphicsState.reverseOrientation;
        } else if (d == "Material") {
            std::string n; ParamSet ps;
            if (!named(&n, &ps)) return false;
            api.graphicsState.material = n;
            api.graphicsState.materialParams = ps;
            api.graphicsState.currentNamedMaterial.clear();
        } else if (d == "Texture") {   // api.cpp:910-936: Texture "name" "float|color|spectrum" "class" parameters
            std::string texName, texType, texClass;
            for (std::string *dst : {&texName, &texType, &texClass}) {
                if (tok.type != Token::STRING) return api.Fail("Texture: expected \"name\" \"type\" \"class\"");
                *dst = tok.text;
                tok = lx.Next();
            }
            ParamSet ps;
            if (!ParseParams(lx, tok, &ps, &perr)) return api.Fail("Texture: " + perr);
            g_texState = &api.graphicsState;
            const bool isFloat = texType == "float";
            if (!isFloat && texType != "color" && texType != "spectrum") {
                fprintf(stderr, "Error: Texture type \"%s\" unknown.\n", texType.c_str());
            } else if (texClass == "constant") {   // textures/constant.cpp:39-48
                if (isFloat) api.graphicsState.floatTextures[texName] = ps.FindOneFloat("value", 1.f);
                else api.graphicsState.spectrumTextures[texName] = ps.FindOneSpectrum("value", V3(1, 1, 1));
            } else if (texClass == "scale") {      // textures/scale.cpp:39-51, scale.h: tex1 * tex2
                if (isFloat) api.graphicsState.floatTextures[texName] = TexFloat(ps, ps, "tex1", 1.f) * TexFloat(ps, ps, "tex2", 1.f);
                else {
                    const V3 a = TexSpectrum(ps, ps, "tex1", V3(1, 1, 1)), b = TexSpectrum(ps, ps, "tex2", V3(1, 1, 1));
                    api.graphicsState.spectrumTextures[texName] = V3(a.x * b.x, a.y * b.y, a.z * b.z);
                }
            } else if (texClass == "mix") {        // textures/mix.cpp:39-54, mix.h: (1 - amt) * t1 + amt * t2
                const float amt = TexFloat(ps, ps, "amount", 0.5f);
                if (isFloat) {
                    const float t1 = TexFloat(ps, ps, "tex1", 0.f), t2 = TexFloat(ps, ps, "tex2", 1.f);
                    api.graphicsState.floatTextures[texName] = (1.f - amt) * t1 + amt * t2;
                } else {
                    const V3 t1 = TexSpectrum(ps, ps, "tex1", V3(0, 0, 0)), t2 = TexSpectrum(ps, ps, "tex2", V3(1, 1, 1));
                    api.graphicsState.spectrumTextures[texName] = V3((1.f - amt) * t1.x + amt * t2.x, (1.f - amt) * t1.y + amt * t2.y,
                                                                     (1.f - amt) * t1.z + amt * t2.z);
                }
            } else
                return api.Fail("texture \"" + texClass + "\" is outside the hot-path scope (constant, scale and mix of constants fold into the material)");
        } else if (d == "MakeNamedMaterial") {   // api.cpp:947-961: the new material's own parameters first, then the current Material's
            std::string n; ParamSet ps;
            if (!named(&n, &ps)) return false;
            const std::string type = ps.FindOneString("type", api.graphicsState.materialParams.FindOneString("type", ""));
            if (type.empty()) fprintf(stderr, "Error: No parameter string \"type\" found in MakeNamedMaterial\n");
            else {
                const int idx = api.MakeMaterialOfType(type, ps, api.graphicsState.materialParams);
                if (idx < 0) return false;
                api.graphicsState.namedMaterials[n] = idx;
            }
        } else if (d == "NamedMaterial") {   // api.cpp:965-968
            if (tok.type != Token::STRING) return api.Fail("NamedMaterial: expected a quoted name");
            api.graphicsState.currentNamedMaterial = tok.text;
            tok = lx.Next();
        } else if (d == "AreaLightSource") {
            std::string n; ParamSet ps;
            if (!named(&n, &ps)) return false;
            api.graphicsState.areaLight = n;
            api.graphicsState.areaLightParams = ps;
        } else if (d == "LightSource") {
            std::string n; ParamSet ps;
            if (!named(&n, &ps)) return false;
            if (!api.DoLight(n, ps)) return false;
        } else if (d == "Shape") {
            std::string n; ParamSet ps;
            if (!named(&n, &ps)) return false;
            if (!api.DoShape(n, ps)) return false;
        } else if (d == "Include") {
            if (tok.type != Token::STRING) return api.Fail("Include: expected filename");
            std::string fn = ResolvePath(api.baseDir, tok.text);
            tok = lx.Next();
            if (!ParseStream(api, fn, depth + 1)) return false;
        } else if (d == "WorldEnd") {
            api.worldEnded = true;
            return true;
        } else {
            return api.Fail("directive \"" + d + "\" is outside the hot-path scope of this front end (" + filename + ":" + std::to_string(lx.line) + ")");
        }
        if (api.worldEnded) return true;
    }
    return true;
}

}  // namespace

bool LoadSceneFile(const std::string &filename, const Options &opt, LoadedScene *out) {
    Api api;
    api.opt = opt;
    api.scene.reset(new Scene);
    size_t slash = filename.rfind('/');
    api.baseDir = slash == std::string::npos ? "" : filename.substr(0, slash);
    g_paramBaseDir = api.baseDir;
    if (!ParseStream(api, filename, 0) || !api.error.empty()) {
        out->error = api.error;
        return false;
    }
    if (!api.worldEnded) { out->error = "scene file has no WorldEnd"; return false; }
    if (!api.Finish(out)) { out->error = api.error; return false; }
    return true;
}

// ---------------------------------------------------------------- synthetic scene (config C5)
// N random small triangles in [-1,1]^3 (pbrt RNG = MT19937, seed 1: per triangle the
// centre c = 2u-1 (3 draws), edges e1, e2 = 0.02 (2u-1) (6 draws), vertices c, c+e1, c+e2, unindexed),
// matte Kd .5, a 20x20 floor quad at y = -1.05, a disk area light L = 40 (radius 1, y = +3, facing -y),
// a point light I = 5 at (0,0,4), camera LookAt 0 0 4.5  0 0 0  0 1 0, fov 40.
// The generator either WRITES the scene as pbrt text (so the unmodified reference can render the same
// scene; used for parity at small N) or builds the Scene directly through the same Api calls the parser
// makes (used for the 10 M-triangle runs, where pbrt's text front end is the bottleneck and its float-typed
// "integer indices" cannot even represent indices > 2^24).  Both routes produce bit-identical floats:
// the text route prints them with %.9g.
namespace {
struct SynthRNG {   // RNG::RandomFloat, core/rng.cpp:43-107
    uint32_t mt[624];
    int mti;
    explicit SynthRNG(uint32_t seed) {
        mt[0] = seed;
        for (mti = 1; mti < 624; mti++) mt[mti] = 1812433253u * (mt[mti - 1] ^ (mt[mti - 1] >> 30)) + (uint32_t)mti;
    }
    uint32_t U() {
        if (mti >= 624) {
            int kk;
            uint32_t y;
            for (kk = 0; kk < 624 - 397; kk++) { y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu); mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u); }
            for (; kk < 623; kk++) { y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu); mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u); }
            y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
            mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            mti = 0;
        }
        uint32_t y = mt[mti++];
        y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
        return y;
    }
    float F() { return (U() & 0xffffff) / float(1 << 24); }
};
const int kSynthTrisPerMesh = 1000000;   // keeps every vertex index below 2^24 (pbrt parses indices as floats)
void SynthTriangle(SynthRNG &rng, V3 v[3]) {
    V3 c;
    c.x = 2.f * rng.F() - 1.f; c.y = 2.f * rng.F() - 1.f; c.z = 2.f * rng.F() - 1.f;
    V3 e1, e2;
    e1.x = 0.02f * (2.f * rng.F() - 1.f); e1.y = 0.02f * (2.f * rng.F() - 1.f); e1.z = 0.02f * (2.f * rng.F() - 1.f);
    e2.x = 0.02f * (2.f * rng.F() - 1.f); e2.y = 0.02f * (2.f * rng.F() - 1.f); e2.z = 0.02f * (2.f * rng.F() - 1.f);
    v[0] = c; v[1] = c + e1; v[2] = c + e2;
}
}  // namespace

bool WriteSyntheticSceneFile(const std::string &filename, const SyntheticParams &sp, std::string *err) {
    FILE *f = fopen(filename.c_str(), "w");
    if (!f) { *err = "cannot write " + filename; return false; }
    fprintf(f, "LookAt 0 0 4.5  0 0 0  0 1 0\nCamera \"perspective\" \"float fov\" [40]\n");
    fprintf(f, "Film \"image\" \"integer xresolution\" [%d] \"integer yresolution\" [%d] \"string filename\" \"%s\"\n",
            sp.xres, sp.yres, sp.outfile.c_str());
    fprintf(f, "Sampler \"lowdiscrepancy\" \"integer pixelsamples\" [%d]\n", sp.spp);
    fprintf(f, "SurfaceIntegrator \"path\" \"integer maxdepth\" [%d]\nWorldBegin\n", sp.maxDepth);
    fprintf(f, "AttributeBegin\nAreaLightSource \"area\" \"color L\" [40 40 40]\nTranslate 0 3 0\nRotate 90 1 0 0\n"
               "Shape \"disk\" \"float radius\" [1]\nAttributeEnd\n");
    fprintf(f, "LightSource \"point\" \"color I\" [5 5 5] \"point from\" [0 0 4]\n");
    fprintf(f, "Material \"matte\" \"color Kd\" [.5 .5 .5]\n");
    fprintf(f, "Shape \"trianglemesh\" \"point P\" [-10 -1.05 -10  10 -1.05 -10  10 -1.05 10  -10 -1.05 10] "
               "\"integer indices\" [0 1 2 2 3 0]\n");
    SynthRNG rng(1);
    for (long long t0 = 0; t0 < sp.nTriangles; t0 += kSynthTrisPerMesh) {
        const int n = (int)std::min<long long>(kSynthTrisPerMesh, sp.nTriangles - t0);
        fprintf(f, "Shape \"trianglemesh\" \"point P\" [\n");
        for (int i = 0; i < n; ++i) {
            V3 v[3];
            SynthTriangle(rng, v);
            fprintf(f, "%.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g\n", v[0].x, v[0].y, v[0].z, v[1].x, v[1].y, v[1].z,
                    v[2].x, v[2].y, v[2].z);
        }
        fprintf(f, "] \"integer indices\" [\n");
        for (int i = 0; i < 3 * n; ++i) fprintf(f, "%d%c", i, (i % 24 == 23) ? '\n' : ' ');
        fprintf(f, "]\n");
    }
    fprintf(f, "WorldEnd\n");
    fclose(f);
    return true;
}

bool LoadSyntheticScene(const SyntheticParams &sp, const Options &opt, LoadedScene *out) {
    Api api;
    api.opt = opt;
    api.scene.reset(new Scene);
    auto I = [](int v) { return std::vector<int>(1, v); };
    auto F1 = [](float v) { return std::vector<float>(1, v); };
    // LookAt / Camera / Film / Sampler / SurfaceIntegrator, exactly as the text route's directives
    api.ForActive([&](Transform &t) { t = t * LookAt(V3(0, 0, 4.5f), V3(0, 0, 0), V3(0, 1, 0), nullptr); });
    api.CameraName = "perspective";
    api.CameraParams.floats["fov"] = F1(40.f);
    api.CameraToWorld = InverseSet(api.curTransform);
    api.namedCoordinateSystems["camera"] = api.CameraToWorld;
    api.FilmParams.ints["xresolution"] = I(sp.xres);
    api.FilmParams.ints["yresolution"] = I(sp.yres);
    api.FilmParams.strings["filename"] = std::vector<std::string>(1, sp.outfile);
    api.SamplerParams.ints["pixelsamples"] = I(sp.spp);
    api.SurfIntegratorName = "path";
    api.SurfIntegratorParams.ints["maxdepth"] = I(sp.maxDepth);
    // WorldBegin
    api.inWorld = true;
    api.curTransform = TransformSet();
    api.namedCoordinateSystems["world"] = api.curTransform;
    {   // disk area light
        api.pushedGraphicsStates.push_back(api.graphicsState);
        api.pushedTransforms.push_back(api.curTransform);
        api.graphicsState.areaLight = "area";
        api.graphicsState.areaLightParams.spectra["L"] = std::vector<V3>(1, V3(40, 40, 40));
        api.ForActive([&](Transform &t) { t = t * Translate(V3(0, 3, 0)); });
        api.ForActive([&](Transform &t) { t = t * Rotate(90, V3(1, 0, 0)); });
        ParamSet ps;
        ps.floats["radius"] = F1(1.f);
        if (!api.DoShape("disk", ps)) { out->error = api.error; return false; }
        api.graphicsState = api.pushedGraphicsStates.back(); api.pushedGraphicsStates.pop_back();
        api.curTransform = api.pushedTransforms.back(); api.pushedTransforms.pop_back();
    }
    {
        ParamSet ps;
        ps.spectra["I"] = std::vector<V3>(1, V3(5, 5, 5));
        ps.points["from"] = std::vector<V3>(1, V3(0, 0, 4));
        if (!api.DoLight("point", ps)) { out->error = api.error; return false; }
    }
    api.graphicsState.material = "matte";
    api.graphicsState.materialParams = ParamSet();
    api.graphicsState.materialParams.spectra["Kd"] = std::vector<V3>(1, V3(.5f, .5f, .5f));
    {
        ParamSet ps;
        ps.points["P"] = {V3(-10, -1.05f, -10), V3(10, -1.05f, -10), V3(10, -1.05f, 10), V3(-10, -1.05f, 10)};
        ps.ints["indices"] = {0, 1, 2, 2, 3, 0};
        if (!api.DoShape("trianglemesh", ps)) { out->error = api.error; return false; }
    }
    SynthRNG rng(1);
    PrepTimer tgen("synthetic triangles + meshes");
    for (long long t0 = 0; t0 < sp.nTriangles; t0 += kSynthTrisPerMesh) {
        const int n = (int)std::min<long long>(kSynthTrisPerMesh, sp.nTriangles - t0);
        ParamSet ps;
        std::vector<V3> &P = ps.points["P"];
        std::vector<int> &idx = ps.ints["indices"];
        P.resize(3 * (size_t)n);
        idx.resize(3 * (size_t)n);
        for (int i = 0; i < n; ++i) {
            V3 v[3];
            SynthTriangle(rng, v);
            P[3 * i] = v[0]; P[3 * i + 1] = v[1]; P[3 * i + 2] = v[2];
        }
        for (int i = 0; i < 3 * n; ++i) idx[i] = i;
        if (!api.DoShape("trianglemesh", ps)) { out->error = api.error; return false; }
    }
    tgen.Stop();
    api.worldEnded = true;
    if (!api.Finish(out)) { out->error = api.error; return false; }
    return true;
}

}  // namespace pb
