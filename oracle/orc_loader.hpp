// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_scalar.hpp).
//
// Scene loading for the oracle: description node -> oracle object model, following
// /root/reference/source/rt/scene_loader.d:62-83 (section order), :100-133 (set/setTo), :142-203
// (extractValue / createObject / named entities) and each class's `deserialize`
// (global_settings.d:47-71, camera.d:238-255, light.d:39-43,77-82, geometry.d:61-64,132-140,
// 237-241,339-348, texture.d:56-61,88-94,128-142, shader.d:40-44,137-147,263-280, node.d:70-94).
// BMP texel decode follows /root/reference/source/imageio/bmp.d:60-193,404-422.
// The text parsers are shared with the product host (chess2rt_b200/host/scene_text.hpp); the
// mapping from description nodes to objects below is the oracle's own.
#pragma once
#include <map>

#include "../chess2rt_b200/host/scene_text.hpp"
#include "orc_scene.hpp"

namespace orc {

using c2rt_text::DscNode;
using c2rt_text::ParseError;

struct LoadError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// ---------------------------------------------------------------- imageio/bmp.d
inline uint32_t rd_u32(const std::vector<uint8_t>& b, size_t o) {
    if (o + 4 > b.size()) throw LoadError("BMP: truncated file");
    return uint32_t(b[o]) | (uint32_t(b[o + 1]) << 8) | (uint32_t(b[o + 2]) << 16) | (uint32_t(b[o + 3]) << 24);
}
inline uint16_t rd_u16(const std::vector<uint8_t>& b, size_t o) {
    if (o + 2 > b.size()) throw LoadError("BMP: truncated file");
    return uint16_t(b[o] | (b[o + 1] << 8));
}

// Decodes to packed 0xAARRGGBB words, row-major with y = 0 the LAST scanline stored in the file
// (bmp.d:140 `foreach_reverse (y; 0 .. header.height)`).
inline void decode_bmp(const std::vector<uint8_t>& f, size_t& W, size_t& H, std::vector<uint32_t>& out) {
    if (f.size() < 18 || f[0] != 'B' || f[1] != 'M')
        throw LoadError("Only files beginning with 'BM' are supported!");  // bmp.d:53-55
    uint32_t offsetToPixelArray = rd_u32(f, 10);
    uint32_t ver = rd_u32(f, 14);  // DIB header size selects the version (bmp.d:51,57)
    long width, height;
    unsigned planes, bpp;
    uint32_t colorsUsed = 0;
    size_t paletteElem;
    if (ver == 12) {  // BITMAPCOREHEADER: 16-bit fields, 3-byte palette entries (bmp.d:327-333,404-413)
        width = (int16_t)rd_u16(f, 18);
        height = (int16_t)rd_u16(f, 20);
        planes = rd_u16(f, 22);
        bpp = rd_u16(f, 24);
        paletteElem = 3;
    } else if (ver == 40 || ver == 52 || ver == 56 || ver == 108 || ver == 124) {
        width = (int32_t)rd_u32(f, 18);
        height = (int32_t)rd_u32(f, 22);
        planes = rd_u16(f, 26);
        bpp = rd_u16(f, 28);
        colorsUsed = rd_u32(f, 46);
        paletteElem = 4;
    } else {
        throw LoadError("BMP: unsupported DIB header size " + std::to_string(ver));
    }
    if (planes != 1) throw LoadError("Only .bmp files with 1 color plane are supported.");          // bmp.d:74-77
    if (!(bpp == 1 || bpp == 2 || bpp == 4 || bpp == 8 || bpp == 16 || bpp == 24 || bpp == 32 || bpp == 64))
        throw LoadError("Only .bmp files with 1, 2, 4, 8, 16, 24, 32 or 64 bpp are supported.");    // bmp.d:79-83
    if (width <= 0 || height <= 0) throw LoadError("BMP: non-positive dimensions");
    std::vector<uint32_t> palette;
    if (bpp <= 8) {  // bmp.d:97-110
        uint32_t n = (ver == 12) ? (1u << bpp) : (colorsUsed ? colorsUsed : (1u << bpp));
        size_t po = 14 + ver;
        for (uint32_t i = 0; i < n; i++) {
            size_t o = po + i * paletteElem;
            if (o + paletteElem > f.size()) throw LoadError("BMP: truncated palette");
            uint32_t v = uint32_t(f[o]) | (uint32_t(f[o + 1]) << 8) | (uint32_t(f[o + 2]) << 16);
            if (paletteElem == 4) v |= uint32_t(f[o + 3]) << 24;
            palette.push_back(v);
        }
    }
    W = (size_t)width;
    H = (size_t)height;
    out.assign(W * H, 0);
    size_t pos = offsetToPixelArray;                                  // bmp.d:117
    size_t row_size = bpp / 8 * W;                                    // bmp.d:133
    size_t row_size_padding = ((bpp * W + 31) / 32) * 4;              // bmp.d:134
    if (bpp == 24 || bpp == 32) {
        size_t bytes = bpp / 8;
        for (size_t yy = H; yy-- > 0;) {
            if (pos + W * bytes > f.size()) throw LoadError("BMP: truncated pixel array");
            for (size_t x = 0; x < W; x++) {
                const uint8_t* q = &f[pos + x * bytes];
                uint32_t v = uint32_t(q[0]) | (uint32_t(q[1]) << 8) | (uint32_t(q[2]) << 16);
                if (bytes == 4) v |= uint32_t(q[3]) << 24;
                out[W * yy + x] = v;
            }
            pos += W * bytes + (row_size_padding - row_size);
        }
    } else if (bpp <= 8) {
        // bmp.d:168-187: reads `width` BYTES per row (no padding skip) and unpacks 8/bpp pixels per byte
        size_t maxShift = 8 / bpp;
        uint32_t mask = (1u << bpp) - 1;
        for (size_t yy = H; yy-- > 0;) {
            if (pos + W > f.size()) throw LoadError("BMP: truncated pixel array");
            for (size_t i = 0; i < W; i++) {
                uint8_t pack = f[pos + i];
                for (size_t s = maxShift; s-- > 0;) {
                    uint32_t idx = (pack >> (bpp * s)) & mask;
                    size_t x = i * maxShift + maxShift - (s + 1);
                    if (x >= W) throw LoadError("BMP: packed row overruns the scanline");  // D: range violation
                    if (idx >= palette.size()) throw LoadError("BMP: palette index out of range");
                    out[W * yy + x] = palette[idx];
                }
            }
            pos += W;
        }
    } else {
        throw LoadError("Not implemented: bpp > 8 && bpp != 24 && bpp != 32");  // bmp.d:189-190
    }
}

inline std::vector<uint8_t> read_bytes(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw LoadError("cannot open '" + path + "'");
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

inline void load_bitmap(const std::string& path, Bitmap& bmp) {  // bitmap.d:66-79 + color.d:60-66
    std::vector<uint32_t> px;
    decode_bmp(read_bytes(path), bmp.width, bmp.height, px);
    bmp.px.resize(px.size() * 3);
    const float divider = 1.0f / 255.0f;
    for (size_t i = 0; i < px.size(); i++) {
        bmp.px[3 * i + 0] = float((px[i] >> 16) & 0xff) * divider;
        bmp.px[3 * i + 1] = float((px[i] >> 8) & 0xff) * divider;
        bmp.px[3 * i + 2] = float(px[i] & 0xff) * divider;
    }
}

// ---------------------------------------------------------------- scene_loader.d
struct Loader {
    Scene* scene;
    std::string filePath;
    std::map<std::string, PointLight*> lights;
    std::map<std::string, Geometry*> geometries;
    std::map<std::string, Texture*> textures;
    std::map<std::string, Shader*> shaders;
    std::map<std::string, Node*> nodes;

    // context.set(...) for scalar kinds — returns whether the property was present
    bool set(double& prop, const DscNode& val, const char* name) {
        if (!val.isSpecified(name)) return false;
        prop = val.getChild(name)->getFloat();
        return true;
    }
    bool set(float& prop, const DscNode& val, const char* name) {
        if (!val.isSpecified(name)) return false;
        prop = (float)val.getChild(name)->getFloat();  // to!float(double)
        return true;
    }
    bool set(bool& prop, const DscNode& val, const char* name) {
        if (!val.isSpecified(name)) return false;
        prop = val.getChild(name)->getBool();
        return true;
    }
    bool set(uint32_t& prop, const DscNode& val, const char* name) {
        if (!val.isSpecified(name)) return false;
        long long v = val.getChild(name)->getInt();
        if (v < 0 || v > 0xffffffffll) throw LoadError(std::string("integer out of range for '") + name + "'");  // to!uint throws
        prop = (uint32_t)v;
        return true;
    }
    bool set(size_t& prop, const DscNode& val, const char* name) {
        if (!val.isSpecified(name)) return false;
        long long v = val.getChild(name)->getInt();
        if (v < 0) throw LoadError(std::string("negative integer for '") + name + "'");
        prop = (size_t)v;
        return true;
    }
    bool set(std::string& prop, const DscNode& val, const char* name) {
        if (!val.isSpecified(name)) return false;
        prop = val.getChild(name)->getString();
        return true;
    }
    static void three(const DscNode& n, double out[3]) {  // scene_loader.d:152-157
        auto vals = n.getValues();
        if (vals.size() < 3) throw LoadError("a Vector/Color needs three values");
        for (int i = 0; i < 3; i++) out[i] = vals[i].asDouble();
    }
    bool setVec(double out[3], const DscNode& val, const char* name) {
        if (!val.isSpecified(name)) return false;
        three(*val.getChild(name), out);
        return true;
    }
    bool setColor(Color& c, const DscNode& val, const char* name) {
        if (!val.isSpecified(name)) return false;
        double d[3];
        three(*val.getChild(name), d);
        c = Color::fromFloats((float)d[0], (float)d[1], (float)d[2]);  // Color(float,float,float) narrows the doubles
        return true;
    }
    template <class M>
    static typename M::mapped_type lookup(const M& m, const std::string& key, const char* what) {
        auto it = m.find(key);
        if (it == m.end()) throw LoadError(std::string("unknown ") + what + " '" + key + "'");  // D: RangeError
        return it->second;
    }
    template <class T, class M>
    void registerNamed(const DscNode& n, T* obj, M& m) {  // scene_loader.d:195-200
        if (!n.hasName()) return;
        std::string name = n.getName();
        if (m.count(name)) throw LoadError("entity with duplicate name: " + name);
        m[name] = obj;
    }
    std::string resolveRelativePath(const std::string& p) const {  // scene_loader.d:135-138
        if (!p.empty() && p[0] == '/') return p;
        return c2rt_text::Document::dirName(filePath) + "/" + p;
    }

    void loadSettings(const DscNode& root) {
        GlobalSettings& s = scene->settings;
        if (!root.isSpecified("GlobalSettings")) return;  // new GlobalSettings()
        auto v = root.getChild("GlobalSettings");
        set(s.frameWidth, *v, "frameWidth");
        set(s.frameHeight, *v, "frameHeight");
        set(s.fullscreen, *v, "fullscreen");
        set(s.allowResize, *v, "allowResize");
        set(s.dynamicAspectRatio, *v, "dynamicAspectRatio");
        set(s.interactive, *v, "interactive");
        set(s.bucketSize, *v, "bucketSize");
        set(s.threadCount, *v, "threadCount");
        set(s.prepassEnabled, *v, "prepassEnabled");
        set(s.prepassOnly, *v, "prepassOnly");
        set(s.GIEnabled, *v, "GIEnabled");
        set(s.AAEnabled, *v, "AAEnabled");
        set(s.AAThreshold, *v, "AAThreshold");
        set(s.maxTraceDepth, *v, "maxTraceDepth");
        set(s.pathsPerPixel, *v, "pathsPerPixel");
        setColor(s.ambientLightColor, *v, "ambientLightColor");
        set(s.debugEnabled, *v, "debugEnabled");
    }
    void loadCamera(const DscNode& root) {
        Camera& c = scene->camera;
        if (!root.isSpecified("Camera")) return;  // new Camera(): deserialize never runs, frame size stays 0
        auto v = root.getChild("Camera");
        setVec(c.pos, *v, "pos");
        set(c.yaw, *v, "yaw");
        set(c.pitch, *v, "pitch");
        set(c.roll, *v, "roll");
        set(c.fov, *v, "fov");
        set(c.focalPlaneDist, *v, "focalPlaneDist");
        set(c.fNumber, *v, "fNumber");
        set(c.dof, *v, "dof");
        set(c.numSamples, *v, "numSamples");
        set(c.stereoSeparation, *v, "stereoSeparation");
        c.discMultiplier = 10.0 / c.fNumber;
        c.setFrameSize(scene->settings.frameWidth, scene->settings.frameHeight);
    }
    void loadLights(const DscNode& root) {
        if (!root.isSpecified("Lights")) return;
        for (auto& n : root.getChild("Lights")->getChildren()) {
            if (n->getType() != "PointLight") throw LoadError("Unknown object type (or not yet supported): " + n->getType());
            auto l = std::make_unique<PointLight>();
            setColor(l->lightColor, *n, "color");
            float power = std::numeric_limits<float>::quiet_NaN();
            set(power, *n, "power");
            l->lightPower = mk_colf(power);
            double p[3] = {NAN, NAN, NAN};
            setVec(p, *n, "pos");
            l->pos = Vec3(mk_real(p[0]), mk_real(p[1]), mk_real(p[2]));
            registerNamed(*n, l.get(), lights);
            scene->lights.push_back(std::move(l));
        }
    }
    void loadGeometries(const DscNode& root) {
        if (!root.isSpecified("Geometries")) return;
        for (auto& n : root.getChild("Geometries")->getChildren()) {
            std::string type = n->getType();
            std::unique_ptr<Geometry> g;
            if (type == "Plane") {
                auto p = std::make_unique<Plane>();
                double y = NAN;
                set(y, *n, "y");
                p->y = mk_real(y);
                g = std::move(p);
            } else if (type == "Sphere") {
                auto s = std::make_unique<Sphere>();
                double c[3];
                if (setVec(c, *n, "center")) s->center = Vec3(mk_real(c[0]), mk_real(c[1]), mk_real(c[2]));
                double R;
                if (set(R, *n, "R")) s->R = mk_real(R);
                g = std::move(s);
            } else if (type == "Cube") {
                auto c = std::make_unique<Cube>();
                double ctr[3];
                if (setVec(ctr, *n, "center")) c->center = Vec3(mk_real(ctr[0]), mk_real(ctr[1]), mk_real(ctr[2]));
                double side;
                if (set(side, *n, "side")) c->side = mk_real(side);
                g = std::move(c);
            } else if (type == "CsgUnion" || type == "CsgInter" || type == "CsgDiff") {
                auto c = std::make_unique<CsgOp>(type == "CsgUnion" ? CsgOp::Union : type == "CsgInter" ? CsgOp::Inter : CsgOp::Diff);
                std::string name;
                set(name, *n, "left");
                c->left = lookup(geometries, name, "geometry");
                set(name, *n, "right");  // if absent, `geomName` keeps the left name (geometry.d:341-347)
                c->right = lookup(geometries, name, "geometry");
                g = std::move(c);
            } else {
                throw LoadError("Unknown object type (or not yet supported): " + type);
            }
            registerNamed(*n, g.get(), geometries);
            scene->geometries.push_back(std::move(g));
        }
    }
    void loadTextures(const DscNode& root) {
        if (!root.isSpecified("Textures")) return;
        for (auto& n : root.getChild("Textures")->getChildren()) {
            std::string type = n->getType();
            std::unique_ptr<Texture> t;
            if (type == "Checker") {
                auto c = std::make_unique<Checker>();
                setColor(c->color1, *n, "color1");
                setColor(c->color2, *n, "color2");
                double size;
                if (set(size, *n, "size")) c->size = mk_real(size);
                t = std::move(c);
            } else if (type == "Procedure2") {
                auto p = std::make_unique<Procedure2>();
                auto colors = [&](const char* key, std::vector<Color>& dst) {
                    if (!n->isSpecified(key)) return;
                    for (auto& ch : n->getChild(key)->getChildren()) {  // scene_loader.d:173-174
                        double d[3];
                        three(*ch, d);
                        dst.push_back(Color::fromFloats((float)d[0], (float)d[1], (float)d[2]));
                    }
                };
                auto freqs = [&](const char* key, std::vector<real>& dst) {
                    if (!n->isSpecified(key)) return;
                    for (auto& v : n->getChild(key)->getValues()) dst.push_back(mk_real(v.asDouble()));  // :169-171
                };
                colors("colorU", p->colorU);
                colors("colorV", p->colorV);
                freqs("freqU", p->freqU);
                freqs("freqV", p->freqV);
                if (p->colorU.size() < 3 || p->colorV.size() < 3 || p->freqU.size() < 3 || p->freqV.size() < 3)
                    throw LoadError("Procedure2 needs three colorU/colorV/freqU/freqV entries");  // D: RangeError at render time
                t = std::move(p);
            } else if (type == "BitmapTexture") {
                auto b = std::make_unique<BitmapTexture>();
                set(b->scaling, *n, "scaling");
                set(b->assumedGamma, *n, "assumedGamma");
                std::string file;
                set(file, *n, "file");
                load_bitmap(resolveRelativePath(file), b->bmp);
                if (b->assumedGamma == 2.2f) b->bmp.decompressGamma_sRGB();
                else if (b->assumedGamma != 1 && b->assumedGamma > 0 && b->assumedGamma < 10)
                    b->bmp.decompressGamma(b->assumedGamma);
                t = std::move(b);
            } else {
                throw LoadError("Unknown object type (or not yet supported): " + type);
            }
            registerNamed(*n, t.get(), textures);
            scene->textures.push_back(std::move(t));
        }
    }
    // environment.d:12-14 reads no keys.  EXTENSION (orc_scene.hpp Environment): `folder` names a directory with the six face
    // BMPs, `assumedGamma` is applied at load time exactly like BitmapTexture's (texture.d:137-141).
    void loadEnvironment(const DscNode& root) {
        if (!root.isSpecified("Environment")) return;
        auto n = root.getChild("Environment");
        std::string folder;
        set(folder, *n, "folder");
        if (folder.empty()) return;
        Environment& e = scene->environment;
        set(e.assumedGamma, *n, "assumedGamma");
        for (int f = 0; f < 6; f++) {
            load_bitmap(resolveRelativePath(folder + "/" + Environment::faceName(f) + ".bmp"), e.faces[f]);
            if (e.assumedGamma == 2.2f) e.faces[f].decompressGamma_sRGB();
            else if (e.assumedGamma != 1 && e.assumedGamma > 0 && e.assumedGamma < 10) e.faces[f].decompressGamma(e.assumedGamma);
        }
        e.cubemap = true;
    }
    const Texture* optionalTexture(const DscNode& n) {
        std::string t;
        set(t, n, "texture");
        auto it = textures.find(t);
        return it == textures.end() ? nullptr : it->second;
    }
    void loadShaders(const DscNode& root) {
        if (!root.isSpecified("Shaders")) return;
        for (auto& n : root.getChild("Shaders")->getChildren()) {
            std::string type = n->getType();
            std::unique_ptr<Shader> s;
            if (type == "Lambert") {
                auto l = std::make_unique<Lambert>();
                setColor(l->color, *n, "color");
                l->texture = optionalTexture(*n);
                s = std::move(l);
            } else if (type == "Phong") {
                auto p = std::make_unique<Phong>();
                setColor(p->color, *n, "color");
                double e = raw(p->exponent);
                set(e, *n, "exponent");
                e = e < 1e-6 ? 1e-6 : (e > 1e6 ? 1e6 : e);  // shader.d:268
                p->exponent = mk_real(e);
                float st = raw(p->strength);
                set(st, *n, "strength");
                st = st < 0.f ? 0.f : (st > 1e6f ? 1e6f : st);  // shader.d:271
                p->strength = mk_colf(st);
                p->texture = optionalTexture(*n);
                s = std::move(p);
            } else {
                throw LoadError("Unknown object type (or not yet supported): " + type);
            }
            s->scene = scene;
            registerNamed(*n, s.get(), shaders);
            scene->shaders.push_back(std::move(s));
        }
    }
    void loadNodes(const DscNode& root) {
        if (!root.isSpecified("Nodes")) return;
        for (auto& n : root.getChild("Nodes")->getChildren()) {
            if (n->getType() != "Node") throw LoadError("Unknown object type (or not yet supported): " + n->getType());
            auto node = std::make_unique<Node>();
            std::string geom, shad, bump;
            set(geom, *n, "geometry");
            set(shad, *n, "shader");
            set(bump, *n, "bump");
            node->geom = lookup(geometries, geom, "geometry");
            node->shader = lookup(shaders, shad, "shader");
            auto bt = textures.find(bump);
            node->bumpmap = bt == textures.end() ? nullptr : bt->second;
            double v[3];
            if (setVec(v, *n, "scale")) node->transform.scale(v[0], v[1], v[2]);
            if (setVec(v, *n, "rotate")) node->transform.scale(v[0], v[1], v[2]);  // node.d:89-90: "rotate" calls scale (quirk)
            if (setVec(v, *n, "translate")) node->transform.translate(v[0], v[1], v[2]);
            registerNamed(*n, node.get(), nodes);
            scene->nodes.push_back(std::move(node));
        }
    }
};

inline std::unique_ptr<Scene> parseSceneFromFile(const std::string& filename) {  // scene_loader.d:20-83
    c2rt_text::Document doc(filename);
    auto root = doc.root();
    auto scene = std::make_unique<Scene>();
    Loader L;
    L.scene = scene.get();
    L.filePath = filename;
    if (root->isSpecified("Name")) scene->name = root->getChild("Name")->getString();
    L.loadSettings(*root);
    L.loadCamera(*root);
    L.loadEnvironment(*root);
    L.loadLights(*root);
    L.loadGeometries(*root);
    L.loadTextures(*root);
    L.loadShaders(*root);
    L.loadNodes(*root);
    return scene;
}

}  // namespace orc
