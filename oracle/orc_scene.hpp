// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_scalar.hpp).  PARITY UNPINNED: the reference has no
// test on the render path and no D toolchain exists here, so this restatement cannot be checked
// against reference output; it follows the reference source line by line instead.
//
// Object model of the render path, restated from /root/reference/source/rt/:
//   ray.d:27-70, intersectable.d:6-33, geometry.d:15-403, transform.d:9-86, node.d:7-49,
//   light.d:6-75, texture.d:6-161, bitmap.d:48-63,105-136, shader.d:24-250, color.d:27-229,
//   camera.d:12-174,231-269, global_settings.d:8-35, scene.d:38-78, environment.d:5-10,
//   util/array.d:95-111 (shell sort), util/random.d:19-28 (uniform).
// Geometry in `real` (FP64), colour in `colf` (FP32); every double->float narrowing is where
// the D code narrows.
#pragma once
#include <stdexcept>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "orc_math.hpp"

namespace orc {

// ------------------------------------------------------------------ color.d
struct Color {
    colf r, g, b;
    Color() : r(mk_colf(0.f)), g(mk_colf(0.f)), b(mk_colf(0.f)) {}  // color.d:33 components default to 0
    Color(colf r_, colf g_, colf b_) : r(r_), g(g_), b(b_) {}
    static Color fromFloats(float r_, float g_, float b_) { return Color(mk_colf(r_), mk_colf(g_), mk_colf(b_)); }
    // color.d:60-66 Color(uint): integer channel * (1.0f/255.0f)
    static Color fromRGB(uint32_t rgb) {
        const float divider = 1.0f / 255.0f;
        return fromFloats(float((rgb >> 16) & 0xff) * divider, float((rgb >> 8) & 0xff) * divider,
                          float(rgb & 0xff) * divider);
    }
    void operator+=(const Color& o) { r += o.r; g += o.g; b += o.b; }                  // color.d:91-97
    Color operator+(const Color& o) const { return Color(r + o.r, g + o.g, b + o.b); }  // color.d:122-126
    Color operator*(const Color& o) const { return Color(r * o.r, g * o.g, b * o.b); }
    Color operator*(colf f) const { return Color(r * f, g * f, b * f); }                // color.d:128-132
    Color operator/(colf f) const { return Color(r / f, g / f, b / f); }
    colf intensity() const { return (r + g + b) / mk_colf(3.f); }                       // color.d:141-144
    void adjustSaturation(colf amount) {                                                // color.d:76-82
        colf mid = intensity();
        r = r * amount + mid * (mk_colf(1.f) - amount);
        g = g * amount + mid * (mk_colf(1.f) - amount);
        b = b * amount + mid * (mk_colf(1.f) - amount);
    }
};

// color.d:10-15: anaglyph combination of the left / right eye colours
inline Color combineStereo(Color left, Color right) {
    left.adjustSaturation(mk_colf(0.25f));
    right.adjustSaturation(mk_colf(0.25f));
    return left * Color::fromFloats(1, 0, 0) + right * Color::fromFloats(0, 1, 1);
}

// color.d:194-229: 8-bit sRGB packing through the 4097-entry table, quirks included
// (12.02 linear slope, floor instead of round).
struct SrgbLut {
    uint8_t t[4097];
    SrgbLut() {
        for (int i = 0; i < 4097; i++) t[i] = convert(float(i) / 4096.f);
    }
    static uint8_t roundToByte(float x) { return (uint8_t)std::floor(x * 255.0f); }  // color.d:216-219
    static uint8_t convert(float x) {                                                // color.d:194-207
        if (x <= 0) return 0;
        if (x >= 1) return 255;
        if (x <= 0.0031308f) x = x * 12.02f;
        else x = (float)(1.055 * std::pow((double)x, 1 / 2.4) - 0.055);  // float promoted to double, narrowed on store
        return roundToByte(x);
    }
    uint8_t cached(float x) const {                                                  // color.d:209-214
        if (x <= 0) return 0;
        if (x >= 1) return 255;
        return t[(int)(x * 4096.0f)];
    }
    uint32_t toRGB32(float r, float g, float b) const {                              // color.d:154-162
        return (uint32_t(cached(b)) << 0) | (uint32_t(cached(g)) << 8) | (uint32_t(cached(r)) << 16);
    }
};

// ------------------------------------------------------------------ ray.d / intersectable.d
struct Geometry;

struct Ray {
    Vec3 orig, dir;
    int depth = 0;
    bool diffuse = false;   // RayFlags.Diffuse (ray.d:6-25), set by Lambert.spawnRay
};

struct IntersectionData {
    Vec3 p, normal;
    real dist;
    real u, v;
    const Geometry* g = nullptr;
    Vec3 dNdx, dNdy;
    IntersectionData() : dist(mk_real(std::numeric_limits<double>::quiet_NaN())), u(dist), v(dist) {}
};

inline Vec3 project(const Vec3& v, int a, int b, int c) {   // imported_types.d:44-51
    Vec3 r;
    r[a] = v[0]; r[b] = v[1]; r[c] = v[2];
    return r;
}
inline Vec3 unproject(const Vec3& v, int a, int b, int c) { // imported_types.d:53-60
    Vec3 r;
    r[0] = v[a]; r[1] = v[b]; r[2] = v[c];
    return r;
}
inline Ray project(Ray r, int a, int b, int c) {            // ray.d:64-70
    r.orig = project(r.orig, a, b, c);
    r.dir = project(r.dir, a, b, c);
    return r;
}
inline Vec3 reflect(const Vec3& ray, const Vec3& norm) {    // imported_types.d:62-67
    Vec3 result = ray - (mk_real(2.0) * dot(ray, norm)) * norm;
    normalize(result);
    return result;
}
inline Vec3 faceforward(const Vec3& ray, const Vec3& norm) { // imported_types.d:69-73
    if (dot(ray, norm) < 0) return norm;
    return -norm;
}

// ------------------------------------------------------------------ geometry.d
struct Stats {
    uint64_t primary = 0, shadow = 0, flops = 0, csg_max_crossings = 0;
    uint64_t bounce = 0;   // GI: continuation rays spawned (renderer.d:452-458)
};
inline Stats& tl_stats() {
    static thread_local Stats s;
    return s;
}
// Conditioning probe for the parity tests (orc_capi.cpp orc_pixel_diag): while `on`, the walk records how far the
// farthest camera-ray hit of a pixel lies and how close the pixel's decisions came to a tie — two node candidates, two
// CSG crossings, an occluder against the light's distance, a checker cell edge (relative gaps).  Off in every timed path.
struct Diag {
    bool on = false;
    double max_dist = 0, min_gap = 1e300;
    void gap(double g) { if (g < min_gap) min_gap = g; }
};
inline Diag& tl_diag() {
    static thread_local Diag d;
    return d;
}

struct Geometry {
    virtual ~Geometry() = default;
    virtual bool intersect(const Ray& ray, IntersectionData& data) const = 0;
    virtual bool isInside(const Vec3& p) const = 0;
};

struct Plane final : Geometry {  // geometry.d:15-59
    real y, limit;              // `limit` cannot be loaded from a scene file -> stays NaN -> unbounded
    Plane() : y(mk_real(std::numeric_limits<double>::quiet_NaN())), limit(y) {}
    bool isInside(const Vec3&) const override { return false; }
    bool intersect(const Ray& ray, IntersectionData& data) const override {
        if ((ray.orig.y > y && ray.dir.y > -1e-9) || (ray.orig.y < y && ray.dir.y < 1e-9)) return false;
        real yDiff = ray.dir.y;
        real wantYDiff = ray.orig.y - y;
        real mult = wantYDiff / -yDiff;
        if (mult > data.dist) return false;
        Vec3 p = ray.orig + ray.dir * mult;
        if (r_fabs(p.x) > limit || r_fabs(p.z) > limit) return false;
        data.p = p;
        data.dist = mult;
        data.normal = Vec3(mk_real(0), mk_real(1), mk_real(0));
        data.dNdx = Vec3(mk_real(1), mk_real(0), mk_real(0));
        data.dNdy = Vec3(mk_real(0), mk_real(0), mk_real(1));
        data.u = data.p.x;
        data.v = data.p.z;
        data.g = this;
        return true;
    }
};

// D's PI is an 80-bit `real`; expressions mixing it with doubles are evaluated in 80-bit and
// rounded once on the store to a double (geometry.d:119-121).
static const long double PI_L = 3.14159265358979323846264338327950288L;

struct Sphere final : Geometry {  // geometry.d:73-130
    Vec3 center;
    real R;
    Sphere() : center(mk_real(0), mk_real(0), mk_real(0)), R(mk_real(1)) {}
    bool intersect(const Ray& ray, IntersectionData& info) const override {
        Vec3 H = ray.orig - center;
        real A = sqlen(ray.dir);
        real B = mk_real(2.0) * dot(H, ray.dir);
        real C = sqlen(H) - R * R;
        real Dscr = B * B - mk_real(4.0) * A * C;
        if (Dscr < 0) return false;
        real x1 = (-B + r_sqrt(Dscr)) / (mk_real(2.0) * A);
        real x2 = (-B - r_sqrt(Dscr)) / (mk_real(2.0) * A);
        real sol = x2;
        if (sol < 0) sol = x1;
        if (sol < 0) return false;
        if (sol > info.dist) return false;
        info.dist = sol;
        info.p = ray.orig + ray.dir * sol;
        info.normal = info.p - center;
        normalize(info.normal);
        real angle = r_atan2(info.p.z - center.z, info.p.x - center.x);
        // (PI + angle)/(2*PI) and 1.0 - (PI/2 + asin(..))/PI in 80-bit, one rounding each
        orc_flops(2);
        info.u = mk_real((double)((PI_L + (long double)raw(angle)) / (2 * PI_L)));
        real as = r_asin((info.p.y - center.y) / R);
        orc_flops(3);
        info.v = mk_real((double)(1.0L - (PI_L / 2 + (long double)raw(as)) / PI_L));
        orc_flops(2);
        long double a2 = (long double)raw(angle) + PI_L / 2;
        info.dNdx = Vec3(r_cos(mk_real((double)a2)), mk_real(0), r_sin(mk_real((double)a2)));
        info.dNdy = cross(info.dNdx, info.normal);
        info.g = this;
        return true;
    }
    bool isInside(const Vec3& p) const override { return sqlen(center - p) < R * R; }
};

struct Cube final : Geometry {  // geometry.d:149-235
    Vec3 center;
    real side;
    Cube() : center(mk_real(0), mk_real(0), mk_real(0)), side(mk_real(1)) {}
    bool isInside(const Vec3& p) const override {
        return r_fabs(p.x - center.x) <= side * 0.5 && r_fabs(p.y - center.y) <= side * 0.5 &&
               r_fabs(p.z - center.z) <= side * 0.5;
    }
    bool intersect(const Ray& ray, IntersectionData& data) const override {
        bool found = intersectCubeSide(ray, center, data);
        if (intersectCubeSide(project(ray, 1, 0, 2), project(center, 1, 0, 2), data)) {
            found = true;
            data.normal = unproject(data.normal, 1, 0, 2);
            data.p = unproject(data.p, 1, 0, 2);
        }
        if (intersectCubeSide(project(ray, 0, 2, 1), project(center, 0, 2, 1), data)) {
            found = true;
            data.normal = unproject(data.normal, 0, 2, 1);
            data.p = unproject(data.p, 0, 2, 1);
        }
        if (found) data.g = this;
        return found;
    }

private:
    bool intersectCubeSide(const Ray& ray, const Vec3& center_, IntersectionData& data) const {
        if (r_fabs(ray.dir.y) < 1e-9) return false;
        real halfSide = side * 0.5;
        bool found = false;
        for (int s = -1; s <= 1; s += 2) {
            real yDiff = ray.dir.y;
            real wantYDiff = ray.orig.y - (center_.y + mk_real((double)s) * halfSide);
            real mult = wantYDiff / -yDiff;
            if (mult < 0) continue;
            if (mult > data.dist) continue;
            Vec3 p = ray.orig + ray.dir * mult;
            if (p.x < center_.x - halfSide || p.x > center_.x + halfSide || p.z < center_.z - halfSide ||
                p.z > center_.z + halfSide)
                continue;
            data.p = ray.orig + ray.dir * mult;
            data.dist = mult;
            data.normal = Vec3(mk_real(0), mk_real((double)s), mk_real(0));
            data.dNdx = Vec3(mk_real(1), mk_real(0), mk_real(0));
            data.dNdy = Vec3(mk_real(0), mk_real(0), mk_real((double)s));
            data.u = data.p.x - center_.x;
            data.v = data.p.z - center_.z;
            found = true;
        }
        return found;
    }
};

// util/array.d:95-111 — shell sort, `foreach (ref i, elem; arr)` with the index modified in
// the body, gap sequence n/2 -> (inc==2 ? 1 : int(inc*5.0/11)).  Compared with opCmp on dist
// (intersectable.d:27-32).
inline void shell_sort(std::vector<IntersectionData>& arr) {
    size_t inc = arr.size() / 2;
    while (inc) {
        for (size_t key = 0; key < arr.size(); key++) {
            size_t i = key;
            IntersectionData elem = arr[i];
            while (i >= inc && arr[i - inc].dist > elem.dist) {
                arr[i] = arr[i - inc];
                i -= inc;
            }
            arr[i] = elem;
            key = i;  // the D loop index is a `ref` to the hidden counter
        }
        inc = (inc == 2) ? 1 : (size_t)(int)((double)inc * 5.0 / 11);
    }
}

struct CsgOp : Geometry {  // geometry.d:250-337
    enum Op { Union, Inter, Diff } op;
    const Geometry* left = nullptr;
    const Geometry* right = nullptr;
    explicit CsgOp(Op o) : op(o) {}
    bool boolOp(bool inL, bool inR) const {
        switch (op) {
            case Union: return inL || inR;   // :361-364
            case Inter: return inL && inR;   // :371-374
            default: return inL && !inR;     // :399-402
        }
    }
    static void findAllIntersections(const Geometry* geom, Ray ray, std::vector<IntersectionData>& l) {
        real currentLength = mk_real(0);
        while (true) {
            IntersectionData temp;
            temp.dist = mk_real(1e99);
            if (!geom->intersect(ray, temp)) break;
            temp.dist += currentLength;
            currentLength = temp.dist;
            ray.orig = temp.p + ray.dir * mk_real(1e-6);
            l.push_back(temp);
            if (l.size() > 4096) break;  // safety net only; never reached by convex children
        }
    }
    bool intersectGeneric(const Ray& ray, IntersectionData& data) const {
        std::vector<IntersectionData> leftData, rightData, allData;
        findAllIntersections(left, ray, leftData);
        findAllIntersections(right, ray, rightData);
        for (auto& e : leftData) allData.push_back(e);
        for (auto& e : rightData) allData.push_back(e);
        uint64_t mx = std::max(leftData.size(), rightData.size());
        if (mx > tl_stats().csg_max_crossings) tl_stats().csg_max_crossings = mx;
        shell_sort(allData);
        bool inL = leftData.size() % 2 == 1;
        bool inR = rightData.size() % 2 == 1;
        size_t walked = 0;
        for (auto& current : allData) {
            if (tl_diag().on) {   // gap to the next crossing of the sorted list (a tie would swap the walk order)
                if (walked + 1 < allData.size()) {
                    const double a = raw(current.dist), b = raw(allData[walked + 1].dist);
                    tl_diag().gap(std::fabs(b - a) / std::max(1.0, std::fabs(a)));
                }
                walked++;
            }
            if (current.g == left) inL = !inL;
            else inR = !inR;
            if (boolOp(inL, inR)) {
                if (tl_diag().on) tl_diag().gap(std::fabs(raw(current.dist) - raw(data.dist)) / std::max(1.0, std::fabs(raw(current.dist))));
                if (current.dist > data.dist) return false;
                data = current;
                return true;
            }
        }
        return false;
    }
    bool intersect(const Ray& ray, IntersectionData& data) const override {
        if (op != Diff) return intersectGeneric(ray, data);
        // CsgDiff.intersect geometry.d:382-397
        if (!intersectGeneric(ray, data)) return false;
        if (right->isInside(data.p - ray.dir * mk_real(1e-6)) != right->isInside(data.p + ray.dir * mk_real(1e-6)))
            data.normal = -data.normal;
        return true;
    }
    bool isInside(const Vec3& p) const override { return boolOp(left->isInside(p), right->isInside(p)); }
};

// ------------------------------------------------------------------ transform.d
struct Transform {
    Mat3d transform, inverseTransform, transposedInverse;  // host-side (load-time) doubles
    double offset[3];
    Mat3 T, Ti, TiT;                                        // copies in the (possibly counted) scalar
    Vec3 off;
    void sync() {
        T = to_counted(transform);
        Ti = to_counted(inverseTransform);
        TiT = to_counted(transposedInverse);
        off = Vec3(mk_real(offset[0]), mk_real(offset[1]), mk_real(offset[2]));
    }
    void reset() {  // :24-30
        transform = m3_identity();
        inverseTransform = m3_inverse(transform);
        transposedInverse = m3_transposed(inverseTransform);
        offset[0] = offset[1] = offset[2] = 0.0;
        sync();
    }
    void scale(double x, double y, double z) {  // :32-39
        Mat3d s{};
        s.c[0][0] = x; s.c[1][1] = y; s.c[2][2] = z;
        transform = m3_mul(transform, s);
        inverseTransform = m3_inverse(transform);
        transposedInverse = m3_transposed(inverseTransform);
        sync();
    }
    void rotate(double yaw, double pitch, double roll) {  // :41-50 (never reached from a scene file: node.d:89-90)
        transform = m3_mul(m3_mul(m3_mul(transform, m3_rotate_x(d_radians(pitch))), m3_rotate_y(d_radians(yaw))),
                           m3_rotate_z(d_radians(roll)));
        inverseTransform = m3_inverse(transform);
        transposedInverse = m3_transposed(inverseTransform);
        sync();
    }
    void translate(double x, double y, double z) {  // :52-55
        offset[0] = x; offset[1] = y; offset[2] = z;
        sync();
    }
    Vec3 point(Vec3 P) const { P = mul(P, T); P = P + off; return P; }        // :57-63
    Vec3 undoPoint(Vec3 P) const { P = P - off; P = mul(P, Ti); return P; }   // :65-71
    Vec3 direction(const Vec3& d) const { return mul(d, T); }                 // :73-76
    Vec3 normal(const Vec3& d) const { return mul(d, TiT); }                  // :78-81
    Vec3 undoDirection(const Vec3& d) const { return mul(d, Ti); }            // :83-86
};

// ------------------------------------------------------------------ texture.d / bitmap.d
struct Texture {
    virtual ~Texture() = default;
    virtual Color getTexColor(const Ray& ray, real u, real v, Vec3& normal) const = 0;
};

// x86 cvttsd2si semantics for cast(int) of an out-of-range / NaN double (texture.d:48-49).
inline int32_t d_cast_int(double x) {
    if (!(x > -2147483649.0 && x < 2147483648.0)) return INT32_MIN;
    return (int32_t)x;
}

struct Checker final : Texture {  // texture.d:20-54
    Color color1, color2;
    real size;
    Checker() : color1(Color::fromFloats(0, 0, 0)), color2(Color::fromFloats(1, 1, 1)), size(mk_real(1.0)) {}
    Color getTexColor(const Ray&, real u, real v, Vec3&) const override {
        if (tl_diag().on) {   // distance to the nearest cell edge, in cells
            const double a = raw(u / size), b = raw(v / size);
            tl_diag().gap(std::fabs(a - std::nearbyint(a)));
            tl_diag().gap(std::fabs(b - std::nearbyint(b)));
        }
        int32_t x = d_cast_int(raw(r_floor(u / size)));
        int32_t y = d_cast_int(raw(r_floor(v / size)));
        int32_t white = (int32_t)((uint32_t)x + (uint32_t)y) % 2;
        return white ? color2 : color1;
    }
};

struct Procedure2 final : Texture {  // texture.d:70-86
    std::vector<Color> colorU, colorV;
    std::vector<real> freqU, freqV;
    Color getTexColor(const Ray&, real u, real v, Vec3&) const override {
        Color result = Color::fromFloats(0, 0, 0);
        for (int i = 0; i < 3; i++)
            result += colorU[i] * narrow(r_sin(u * freqU[i])) + colorV[i] * narrow(r_sin(v * freqV[i]));
        return result;
    }
};

struct Bitmap {  // bitmap.d:11-63 over imageio/image.d:18-60
    size_t width = 0, height = 0;
    std::vector<float> px;  // r,g,b per texel, row-major, top row first
    bool empty() const { return px.empty(); }
    Color at(size_t x, size_t y) const {
        const float* q = &px[(width * y + x) * 3];
        return Color::fromFloats(q[0], q[1], q[2]);
    }
    bool isInvalidPos(size_t x, size_t y) const { return empty() || x >= width || y >= height; }
    Color getFilteredPixel(colf x, colf y) const {  // bitmap.d:48-63
        float xr = raw(x), yr = raw(y);
        // cast(size_t) of NaN / negative floats is out of range -> red
        if (!(xr >= 0.f) || !(yr >= 0.f) || isInvalidPos((size_t)xr, (size_t)yr)) return Color::fromFloats(1, 0, 0);
        size_t tx = (size_t)raw(f_floor(x));
        size_t ty = (size_t)raw(f_floor(y));
        size_t tx_next = (tx + 1) % width;
        size_t ty_next = (ty + 1) % height;
        colf p = x - mk_colf((float)tx);
        colf q = y - mk_colf((float)ty);
        colf one = mk_colf(1.0f);
        return at(tx, ty) * ((one - p) * (one - q)) + at(tx_next, ty) * (p * (one - q)) +
               at(tx, ty_next) * ((one - p) * q) + at(tx_next, ty_next) * (p * q);
    }
    void remapRGB(float (*fn)(float, float), float arg) {
        for (auto& c : px) c = fn(c, arg);
    }
    // bitmap.d:116-126.  `^^` on floats is std.math.pow evaluated in `real`, stored to float.
    static float srgbDecode(float x, float) {
        if (x == 0) return 0.0f;
        if (x == 1) return 1.0f;
        if (x <= 0.04045f) return x / 12.92f;
        return (float)powl((long double)((x + 0.055f) / 1.055f), (long double)2.4f);
    }
    static float gammaDecode(float x, float gamma) {  // bitmap.d:129-136
        if (x == 0) return 0.0f;
        if (x == 1) return 1.0f;
        return (float)powl((long double)x, (long double)gamma);
    }
    void decompressGamma_sRGB() { remapRGB(srgbDecode, 0.f); }
    void decompressGamma(float gamma) { remapRGB(gammaDecode, gamma); }
};

struct BitmapTexture final : Texture {  // texture.d:103-161
    Bitmap bmp;
    float scaling = 1;
    float assumedGamma = 2.2f;
    Color getTexColor(const Ray&, real u, real v, Vec3&) const override {
        u *= mk_real((double)scaling);
        v *= mk_real((double)scaling);
        u = u - r_floor(u);
        v = v - r_floor(v);
        colf tx = narrow(u) * mk_colf((float)bmp.width);
        colf ty = narrow(v) * mk_colf((float)bmp.height);
        return bmp.getFilteredPixel(tx, ty);
    }
};

// ------------------------------------------------------------------ environment.d
// environment.d:5-15: the reference's Environment is a stub that returns black, and no cubemap class, loader key or asset
// exists in it.  EXTENSION, PARITY UNPINNED: the same class with an optional `folder` key holds six BMP faces
// (posx, negx, posy, negy, posz, negz .bmp; load-time gamma like BitmapTexture, texture.d:137-141) and getEnvironment(dir)
// returns the bilinear sample of the face the direction's largest component points at — the face / uv convention of the
// ancestor project the README credits (README.md:67).  Oracle and kernel define this together; nothing in the reference pins it.
struct Environment {
    enum Face { PosX, NegX, PosY, NegY, PosZ, NegZ };
    Bitmap faces[6];
    bool cubemap = false;
    float assumedGamma = 2.2f;
    static const char* faceName(int f) {
        static const char* n[6] = {"posx", "negx", "posy", "negy", "posz", "negz"};
        return n[f];
    }
    Color getSide(const Bitmap& bmp, real x, real y) const {   // (x, y) in [-1, 1] -> texel coordinates in [0, size - 1]
        colf tx = narrow((x + mk_real(1.0)) * mk_real(0.5) * mk_real((double)(bmp.width - 1)));
        colf ty = narrow((y + mk_real(1.0)) * mk_real(0.5) * mk_real((double)(bmp.height - 1)));
        return bmp.getFilteredPixel(tx, ty);
    }
    Color getEnvironment(const Vec3& dir) const {
        if (!cubemap) return Color::fromFloats(0, 0, 0);   // environment.d:7-10
        const real ax = r_fabs(dir.x), ay = r_fabs(dir.y), az = r_fabs(dir.z);
        // the face of the largest component (x wins ties over y over z), the other two divided by it
        if (ax >= ay && ax >= az) {
            if (!(raw(ax) > 0)) return Color::fromFloats(0, 0, 0);   // zero / NaN direction
            const real vy = dir.y / ax, vz = dir.z / ax;
            return raw(dir.x) < 0 ? getSide(faces[NegX], vz, -vy) : getSide(faces[PosX], -vz, -vy);
        }
        if (ay >= az) {
            const real vx = dir.x / ay, vz = dir.z / ay;
            return raw(dir.y) < 0 ? getSide(faces[NegY], vx, -vz) : getSide(faces[PosY], vx, vz);
        }
        const real vx = dir.x / az, vy = dir.y / az;
        return raw(dir.z) < 0 ? getSide(faces[NegZ], -vx, -vy) : getSide(faces[PosZ], vx, -vy);
    }
};

// ------------------------------------------------------------------ light.d
struct PointLight {
    Vec3 pos;
    Color lightColor;
    colf lightPower = mk_colf(std::numeric_limits<float>::quiet_NaN());
    Color color() const { return lightColor * lightPower; }  // light.d:11-14
    size_t getNumSamples() const { return 1; }               // :56-59
    float solidAngle(const Vec3&) const { return 0.f; }      // :72-75
    void getNthSample(size_t, const Vec3&, Vec3& samplePos, Color& c) const {  // :61-65
        samplePos = pos;
        c = color();
    }
};

// ------------------------------------------------------------------ global_settings.d / camera.d
struct GlobalSettings {
    uint32_t frameWidth = 640, frameHeight = 480;
    bool fullscreen = false, allowResize = false, dynamicAspectRatio = false, interactive = false;
    uint32_t bucketSize = 48, threadCount = 0;
    bool prepassEnabled = true, prepassOnly = false, GIEnabled = false, AAEnabled = true;
    double AAThreshold = 0.1;
    uint32_t pathsPerPixel = 40, maxTraceDepth = 4;
    Color ambientLightColor;  // black
    bool debugEnabled = true;
};

// RNG behind util/random.d `uniform`.  Mode 0 = libc rand() as the reference (non-deterministic
// across thread schedules); mode 1 = the pinned counter-based generator shared
// with the CUDA path (include/c2rt.h c2rt_rng_u31): keyed by (seed, pixel, AA tap, DOF sample, draw).
struct RngState {
    int mode = 1;
    uint64_t seed = 0;
    uint32_t px = 0, py = 0, tap = 0, sample = 0, draw = 0;
};
inline RngState& tl_rng() {
    static thread_local RngState s;
    return s;
}
inline uint32_t rng_u31(uint64_t seed, uint32_t px, uint32_t py, uint32_t tap, uint32_t sample, uint32_t draw) {
    // counter-based: four rounds of the 32-bit "lowbias32" finaliser over (seed, pixel) / (tap, sample) / draw.
    // The pixel and sample rounds do not depend on `draw`, so a compiler hoists them out of the draw sequence.
    uint32_t h = (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x9E3779B9u) ^ (px * 0x85EBCA6Bu);
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    h ^= py * 0xC2B2AE35u;
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    h ^= tap * 0x27D4EB2Fu + sample * 0x165667B1u;
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    h += draw * 0x9E3779B9u;
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    return h >> 1;
}
inline real uniform01() {  // util/random.d:19-28 with a=0, b=1:  0 + (r / RAND_MAX) * 1
    RngState& s = tl_rng();
    double r;
    if (s.mode == 0) r = (double)rand();
    else r = (double)rng_u31(s.seed, s.px, s.py, s.tap, s.sample, s.draw++);
    orc_flops(3);
    return mk_real(0.0 + (r / 2147483647.0) * 1.0);
}

inline void unitDiscSample(real& x, real& y) {  // camera.d:258-269
    orc_flops(2);
    real angle = mk_real((double)((long double)raw(uniform01()) * 2 * PI_L));
    real rad = r_sqrt(uniform01());
    x = r_sin(angle) * rad;
    y = r_cos(angle) * rad;
}

struct Camera {
    size_t frameWidth = 0, frameHeight = 0;
    double aspect = 1.0;
    double pos[3] = {NAN, NAN, NAN};
    double yaw = 0, pitch = 0, roll = 0, fov = 0;
    double focalPlaneDist = 1.0, fNumber = 1.0, discMultiplier = NAN;
    bool dof = false;
    size_t numSamples = 25;
    double stereoSeparation = 0.0;
    // computed by beginFrame (plain doubles: per-frame host work, not counted)
    double upLeft[3], upRight[3], downLeft[3], frontDir[3], rightDir[3], upDir[3];

    void setFrameSize(uint32_t w, uint32_t h) {  // camera.d:231-236
        frameWidth = w;
        frameHeight = h;
        aspect = double(frameWidth) / double(frameHeight);
    }
    void beginFrame() {  // camera.d:77-117
        double x = -aspect, y = +1;
        double cx = x - 0, cy = y - 0, cz = 1.0 - 1.0;
        double lenXY = std::sqrt(0 + cx * cx + cy * cy + cz * cz);
        double wantedLength = (double)tanl((long double)d_radians(fov / 2));
        double scaling = wantedLength / lenXY;
        x *= scaling;
        y *= scaling;
        double ul[3] = {x, y, 1}, ur[3] = {-x, y, 1}, dl[3] = {x, -y, 1};
        Mat3d rotation = m3_mul(m3_mul(m3_rotate_z(d_radians(roll)), m3_rotate_x(d_radians(pitch))),
                                m3_rotate_y(d_radians(yaw)));
        d_mul(ul, rotation, upLeft);
        d_mul(ur, rotation, upRight);
        d_mul(dl, rotation, downLeft);
        double ex[3] = {1, 0, 0}, ey[3] = {0, 1, 0}, ez[3] = {0, 0, 1};
        d_mul(ex, rotation, rightDir);
        d_mul(ey, rotation, upDir);
        d_mul(ez, rotation, frontDir);
        for (int i = 0; i < 3; i++) {
            upLeft[i] += pos[i];
            upRight[i] += pos[i];
            downLeft[i] += pos[i];
        }
    }
    static Vec3 v(const double a[3]) { return Vec3(mk_real(a[0]), mk_real(a[1]), mk_real(a[2])); }

    // camera.d:123-174.  eye: 0 = Stereo3DOffset.None, -1 = Left, +1 = Right
    Ray getScreenRay(real x, real y, int eye = 0) const {
        Ray result;
        Vec3 P = v(pos), UL = v(upLeft), UR = v(upRight), DL = v(downLeft);
        result.orig = P;
        Vec3 target = UL + (UR - UL) * (x / mk_real((double)frameWidth)) + (DL - UL) * (y / mk_real((double)frameHeight));
        result.dir = target - P;
        normalize(result.dir);
        Vec3 right = v(rightDir);
        const real sep = mk_real(eye > 0 ? +stereoSeparation : -stereoSeparation);
        if (eye != 0) result.orig = result.orig + right * sep;   // camera.d:149-152
        if (!dof) return result;
        Vec3 front = v(frontDir), up = v(upDir);
        real cosTheta = dot(result.dir, front);
        real M = mk_real(focalPlaneDist) / cosTheta;
        Vec3 T = result.orig + result.dir * M;
        real dx, dy;
        unitDiscSample(dx, dy);
        dx *= mk_real(discMultiplier);
        dy *= mk_real(discMultiplier);
        result.orig = P + dx * right + dy * up;
        if (eye != 0) result.orig = result.orig + right * sep;   // camera.d:168-170
        result.dir = T - result.orig;
        normalize(result.dir);
        return result;
    }
};

// ------------------------------------------------------------------ shader.d / node.d / scene.d
struct Scene;

struct Shader {
    Color color;  // Shader.color has Color's default (0,0,0) until the ctor / deserialize sets it
    const Scene* scene = nullptr;
    virtual ~Shader() = default;
    virtual Color shade(const Ray& ray, const IntersectionData& data) const = 0;
    // GI branch (shader.d:107-135 Lambert; :252-262 Phong = assert(0), which D keeps in release builds: the reference halts)
    virtual Color eval(const IntersectionData& x, const Ray& w_in, const Ray& w_out) const = 0;
    virtual void spawnRay(const IntersectionData& x, const Ray& w_in, Ray& w_out, Color& colorEval, float& pdf) const = 0;
};

struct ReferenceHalts : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct Node {  // node.d:7-49
    const Geometry* geom = nullptr;
    const Shader* shader = nullptr;
    const Texture* bumpmap = nullptr;  // modifyNormal is a no-op (texture.d:10-12)
    Transform transform;
    Node() { transform.reset(); }
    bool intersect(const Ray& ray, IntersectionData& data) const {
        Ray rayCanonic;
        rayCanonic.orig = transform.undoPoint(ray.orig);
        rayCanonic.dir = transform.undoDirection(ray.dir);
        rayCanonic.depth = ray.depth;
        real oldDist = data.dist;
        real rayDirLength = length(rayCanonic.dir);
        data.dist *= rayDirLength;
        normalize(rayCanonic.dir);
        if (!geom->intersect(rayCanonic, data)) {
            data.dist = oldDist;
            return false;
        }
        data.normal = normalized(transform.normal(data.normal));
        data.dNdx = normalized(transform.direction(data.dNdx));
        data.dNdy = normalized(transform.direction(data.dNdy));
        data.p = transform.point(data.p);
        data.dist /= rayDirLength;
        return true;
    }
};

struct Scene {  // scene.d:38-78
    std::string name;
    GlobalSettings settings;
    Camera camera;
    Environment environment;
    std::vector<std::unique_ptr<PointLight>> lights;
    std::vector<std::unique_ptr<Geometry>> geometries;
    std::vector<std::unique_ptr<Texture>> textures;
    std::vector<std::unique_ptr<Shader>> shaders;
    std::vector<std::unique_ptr<Node>> nodes;

    void beginFrame() { camera.beginFrame(); }

    bool testVisibility(const Vec3& from, const Vec3& to) const {  // scene.d:62-78
        tl_stats().shadow++;
        Ray ray;
        ray.orig = from;
        ray.dir = to - from;
        normalize(ray.dir);
        IntersectionData temp;
        temp.dist = length(to - from);
        if (tl_diag().on) {   // an occluder candidate right at the light's distance
            const double maxd = raw(temp.dist);
            for (auto& node : nodes) {
                IntersectionData probe;
                probe.dist = mk_real(1e99);
                if (node->intersect(ray, probe)) tl_diag().gap(std::fabs(raw(probe.dist) - maxd) / std::max(1.0, maxd));
            }
        }
        for (auto& node : nodes)
            if (node->intersect(ray, temp)) return false;
        return true;
    }
};

// util/random.d:12-28 for integer arguments: a + (r / RAND_MAX) * (b - a), truncated.  r == RAND_MAX yields b, one past
// the range (an out-of-bounds index in the reference): clamped here.
inline size_t uniformIndex(size_t n) {
    double r = raw(uniform01());
    size_t i = (size_t)(0 + r * (double)n);
    return i < n ? i : n - 1;
}

inline Vec3 hemisphereSample(const Vec3& normal) {  // shader.d:156-174
    real u = uniform01(), v = uniform01();
    real theta = mk_real((double)(2 * PI_L * (long double)raw(u)));
    real phi = mk_real((double)((long double)std::acos(2 * raw(v) - 1) - PI_L / 2));
    orc_flops(5);
    Vec3 res(r_cos(theta) * r_cos(phi), r_sin(phi), r_sin(theta) * r_cos(phi));
    if (dot(res, normal) < 0) res = -res;
    return res;
}

struct Lambert final : Shader {  // shader.d:54-135
    const Texture* texture = nullptr;
    Lambert() { color = Color::fromFloats(1, 1, 1); }
    Color shade(const Ray& ray, const IntersectionData& data) const override {
        Vec3 N = faceforward(ray.dir, data.normal);
        Color diffuseColor = texture ? texture->getTexColor(ray, data.u, data.v, N) : color;
        Color lightContrib = scene->settings.ambientLightColor;
        for (auto& light : scene->lights) {
            Color avgColor = Color::fromFloats(0, 0, 0);
            for (size_t j = 0; j < light->getNumSamples(); j++) {
                Vec3 lightPos;
                Color lightColor;
                light->getNthSample(j, data.p, lightPos, lightColor);
                if (raw(lightColor.intensity()) != 0 && scene->testVisibility(data.p + N * mk_real(1e-6), lightPos)) {
                    Vec3 lightDir = lightPos - data.p;
                    normalize(lightDir);
                    real cosTheta = dot(lightDir, N);
                    if (cosTheta > 0) avgColor += lightColor / narrow(sqlen(data.p - lightPos)) * narrow(cosTheta);
                }
            }
            lightContrib += avgColor / mk_colf((float)light->getNumSamples());
        }
        return diffuseColor * lightContrib;
    }
    Color eval(const IntersectionData& x, const Ray& w_in, const Ray& w_out) const override {  // shader.d:107-116
        Vec3 N = faceforward(w_in.dir, x.normal);
        Color diffuseColor = texture ? texture->getTexColor(w_in, x.u, x.v, N) : color;
        real c = dot(w_out.dir, N);
        return diffuseColor * mk_colf((float)(1 / PI_L)) * narrow(c > 0 ? c : mk_real(0.0));
    }
    void spawnRay(const IntersectionData& x, const Ray& w_in, Ray& w_out, Color& colorEval, float& pdf) const override {  // :118-135
        Vec3 N = faceforward(w_in.dir, x.normal);
        Color diffuseColor = texture ? texture->getTexColor(w_in, x.u, x.v, N) : color;
        w_out = w_in;
        w_out.depth++;
        w_out.orig = x.p + N * mk_real(1e-6);
        w_out.dir = hemisphereSample(N);
        w_out.diffuse = true;
        real c = dot(w_out.dir, N);
        colorEval = diffuseColor * mk_colf((float)(1 / PI_L)) * narrow(c > 0 ? c : mk_real(0.0));
        pdf = (float)(1 / (2 * PI_L));
    }
};

struct Phong final : Shader {  // shader.d:177-250
    const Texture* texture = nullptr;
    real exponent = mk_real(16.0);
    colf strength = mk_colf(1.0f);
    Phong() { color = Color::fromFloats(1, 1, 1); }
    Color shade(const Ray& ray, const IntersectionData& data) const override {
        Vec3 N = faceforward(ray.dir, data.normal);
        Color diffuseColor = color;
        if (texture) diffuseColor = texture->getTexColor(ray, data.u, data.v, N);
        Color lightContrib = scene->settings.ambientLightColor;
        Color specular = Color::fromFloats(0, 0, 0);
        for (auto& light : scene->lights) {
            size_t numSamples = light->getNumSamples();
            Color avgColor = Color::fromFloats(0, 0, 0);
            Color avgSpecular = Color::fromFloats(0, 0, 0);
            for (size_t j = 0; j < numSamples; j++) {
                Vec3 lightPos;
                Color lightColor;
                light->getNthSample(j, data.p, lightPos, lightColor);
                if (raw(lightColor.intensity()) != 0 && scene->testVisibility(data.p + N * mk_real(1e-6), lightPos)) {
                    Vec3 lightDir = lightPos - data.p;
                    normalize(lightDir);
                    real cosTheta = dot(lightDir, N);
                    Color baseLight = lightColor / narrow(sqlen(data.p - lightPos));
                    if (cosTheta > 0) avgColor += baseLight * narrow(cosTheta);
                    Vec3 R = reflect(-lightDir, N);
                    real cosGamma = dot(R, -ray.dir);
                    if (cosGamma > 0) avgSpecular += baseLight * narrow(r_pow(cosGamma, exponent)) * strength;
                }
            }
            lightContrib += avgColor / mk_colf((float)numSamples);
            specular += avgSpecular / mk_colf((float)numSamples);
        }
        return diffuseColor * lightContrib + specular;
    }
    Color eval(const IntersectionData&, const Ray&, const Ray&) const override {  // shader.d:258-262
        throw ReferenceHalts("Phong.eval is assert(0) (shader.d:258-262): the reference halts when a GI path lights a Phong surface");
    }
    void spawnRay(const IntersectionData&, const Ray&, Ray&, Color&, float&) const override {  // shader.d:252-256
        throw ReferenceHalts("Phong.spawnRay is assert(0) (shader.d:252-256): the reference halts when a GI path hits a Phong surface");
    }
};

}  // namespace orc
