// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_scalar.hpp).
//
// C entry points over the CPU restatement, for tests/ (ctypes), __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs.  Built twice by oracle/Makefile:
//   liborc.so        plain double/float          (image oracle, CPU baseline timing)
//   liborc_count.so  -DORC_COUNT_FLOPS           (algorithmic FLOP / ray accounting)
#include <chrono>
#include <cstdio>
#include <cstring>

#include "orc_loader.hpp"
#include "orc_render.hpp"

using namespace orc;

struct orc_scene {
    std::unique_ptr<Scene> scene;
};

struct orc_stats {
    uint64_t primary_rays, shadow_rays, flops, prepass_rays, prepass_flops, csg_max_crossings;
    double seconds;
    uint64_t gi_bounce_rays;   // GI: continuation rays (renderer.d:452-458)
};

static thread_local std::string g_err;

extern "C" {

const char* orc_last_error(void) { return g_err.c_str(); }

orc_scene* orc_scene_load(const char* path) {
    try {
        auto h = new orc_scene;
        h->scene = parseSceneFromFile(path);
        return h;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

void orc_scene_free(orc_scene* s) { delete s; }

// "Resolution override" = edit settings.frameWidth/Height and camera.setFrameSize
// (camera.d:254), as if the scene file had been edited.
void orc_scene_set_frame_size(orc_scene* s, uint32_t w, uint32_t h) {
    s->scene->settings.frameWidth = w;
    s->scene->settings.frameHeight = h;
    s->scene->camera.setFrameSize(w, h);
}
void orc_scene_get_frame_size(const orc_scene* s, uint32_t* w, uint32_t* h) {
    *w = s->scene->settings.frameWidth;
    *h = s->scene->settings.frameHeight;
}
// -1 keeps the scene-file value
void orc_scene_override(orc_scene* s, int aa, int dof, int prepass, int num_samples) {
    if (aa >= 0) s->scene->settings.AAEnabled = aa != 0;
    if (dof >= 0) s->scene->camera.dof = dof != 0;
    if (prepass >= 0) s->scene->settings.prepassEnabled = prepass != 0;
    if (num_samples >= 0) s->scene->camera.numSamples = (size_t)num_samples;
}
void orc_scene_info(const orc_scene* s, int32_t out[8]) {
    out[0] = (int32_t)s->scene->nodes.size();
    out[1] = (int32_t)s->scene->geometries.size();
    out[2] = (int32_t)s->scene->shaders.size();
    out[3] = (int32_t)s->scene->textures.size();
    out[4] = (int32_t)s->scene->lights.size();
    out[5] = s->scene->settings.AAEnabled;
    out[6] = s->scene->camera.dof;
    out[7] = (int32_t)s->scene->camera.numSamples;
}

// Camera.beginFrame vectors: pos, upLeft, upRight, downLeft, rightDir, upDir, frontDir (21 doubles)
void orc_camera_vectors(orc_scene* s, double out[21]) {
    Camera& c = s->scene->camera;
    c.beginFrame();
    const double* src[7] = {c.pos, c.upLeft, c.upRight, c.downLeft, c.rightDir, c.upDir, c.frontDir};
    for (int i = 0; i < 7; i++)
        for (int k = 0; k < 3; k++) out[3 * i + k] = src[i][k];
}

// Full frame, reference pass structure.  threads = 0 -> hardware_concurrency().
int orc_render(orc_scene* s, float* rgb, unsigned threads, int rng_mode, uint64_t seed, orc_stats* st) {
    try {
        s->scene->beginFrame();
        Renderer r(*s->scene, rgb);
        r.rngMode = rng_mode;
        r.seed = seed;
        if (threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());
        Stats fin, pre;
        auto t0 = std::chrono::steady_clock::now();
        r.renderRT(threads, fin, pre);
        auto t1 = std::chrono::steady_clock::now();
        if (st) {
            st->primary_rays = fin.primary;
            st->shadow_rays = fin.shadow;
            st->flops = fin.flops;
            st->prepass_rays = pre.primary + pre.shadow;
            st->prepass_flops = pre.flops;
            st->csg_max_crossings = std::max(fin.csg_max_crossings, pre.csg_max_crossings);
            st->seconds = std::chrono::duration<double>(t1 - t0).count();
            st->gi_bounce_rays = fin.bounce;
        }
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// A window of rows [y0, y1) with the reference's per-pixel arithmetic (all taps of a pixel
// summed in the reference's order); used for bounded CPU-baseline samples and big-frame spot checks.
// `rgb` holds (y1-y0)*W*3 floats.
int orc_render_rows(orc_scene* s, float* rgb, uint32_t y0, uint32_t y1, unsigned threads, int rng_mode, uint64_t seed,
                    orc_stats* st) {
    try {
        Scene& sc = *s->scene;
        sc.beginFrame();
        uint32_t W = sc.settings.frameWidth, H = sc.settings.frameHeight;
        if (y1 > H || y0 > y1) { g_err = "row window out of range"; return -1; }
        // Render through a full-width view whose row 0 is y0: the Renderer indexes out[W*y+x].
        float* base = rgb - (size_t)W * y0 * 3;
        Renderer r(sc, base);
        r.rngMode = rng_mode;
        r.seed = seed;
        if (threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());
        std::vector<Box2i> rows;
        const int BS = (int)sc.settings.bucketSize;
        for (uint32_t y = y0; y < y1; y += BS)
            for (uint32_t x = 0; x < W; x += BS)
                rows.push_back({(int)x, (int)y, (int)std::min(W, x + BS), (int)std::min(y1, y + BS)});
        Stats fin;
        auto t0 = std::chrono::steady_clock::now();
        r.parallelBuckets(rows, threads, fin, [&](int x, int y) { r.renderPixelNoAA(x, y); });
        if (sc.settings.AAEnabled) r.parallelBuckets(rows, threads, fin, [&](int x, int y) { r.renderPixelAA(x, y); });
        auto t1 = std::chrono::steady_clock::now();
        if (st) {
            memset(st, 0, sizeof(*st));
            st->primary_rays = fin.primary;
            st->shadow_rays = fin.shadow;
            st->flops = fin.flops;
            st->csg_max_crossings = fin.csg_max_crossings;
            st->seconds = std::chrono::duration<double>(t1 - t0).count();
        }
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// renderer.d:46-57 renderPixel: one un-antialiased sample at the pixel corner + the hit record.
// hit[0]=node index or -1, hit[1]=dist, hit[2..4]=p, hit[5..7]=normal, hit[8]=u, hit[9]=v
int orc_render_pixel(orc_scene* s, int x, int y, int rng_mode, uint64_t seed, float rgb[3], double hit[10]) {
    try {
        Scene& sc = *s->scene;
        sc.beginFrame();
        std::vector<float> scratch((size_t)sc.settings.frameWidth * sc.settings.frameHeight * 3, 0.f);
        Renderer r(sc, scratch.data());
        r.rngMode = rng_mode;
        r.seed = seed;
        Color c = r.renderPixelNoAA(x, y);
        rgb[0] = raw(c.r); rgb[1] = raw(c.g); rgb[2] = raw(c.b);
        if (hit) {
            Ray ray;
            if (sc.camera.dof) {
                RngState& rs = tl_rng();
                rs.mode = rng_mode; rs.seed = seed; rs.px = x; rs.py = y; rs.tap = 0; rs.sample = 0; rs.draw = 0;
                real jx = mk_real((double)x) + uniform01() * mk_real(1.0);
                real jy = mk_real((double)y) + uniform01() * mk_real(1.0);
                ray = sc.camera.getScreenRay(jx, jy, sc.camera.stereoSeparation != 0 ? -1 : 0);
            } else {
                // with stereo on, the record is the LEFT eye's (the first ray traced: renderer.d:309-311)
                ray = sc.camera.getScreenRay(mk_real((double)x), mk_real((double)y), sc.camera.stereoSeparation != 0 ? -1 : 0);
            }
            IntersectionData data;
            data.dist = mk_real(1e99);
            int closest = -1;
            for (size_t i = 0; i < sc.nodes.size(); i++)
                if (sc.nodes[i]->intersect(ray, data)) closest = (int)i;
            hit[0] = closest;
            hit[1] = raw(data.dist);
            hit[2] = raw(data.p.x); hit[3] = raw(data.p.y); hit[4] = raw(data.p.z);
            hit[5] = raw(data.normal.x); hit[6] = raw(data.normal.y); hit[7] = raw(data.normal.z);
            hit[8] = raw(data.u); hit[9] = raw(data.v);
        }
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// Environment.getEnvironment(dir) (environment.d:7-10 black stub; cubemap EXTENSION in orc_scene.hpp) -> rgb; returns 1 when the
// scene's environment is a cubemap, 0 when it is the reference's black one.
int orc_environment(orc_scene* s, const double dir[3], float rgb[3]) {
    Vec3 d;
    d.x = mk_real(dir[0]); d.y = mk_real(dir[1]); d.z = mk_real(dir[2]);
    Color c = s->scene->environment.getEnvironment(d);
    rgb[0] = raw(c.r); rgb[1] = raw(c.g); rgb[2] = raw(c.b);
    return s->scene->environment.cubemap ? 1 : 0;
}

// Conditioning probe of one pixel for the parity tests: out[0] = farthest camera-ray hit of the pixel, out[1] = smallest
// relative gap to a tie met while rendering it (orc_scene.hpp Diag), out[2] = largest colour change of the ORACLE's own
// pixel when every sample position is shifted by (+-eps, +-eps) pixels, out[3..5] = the unshifted colour.
int orc_pixel_diag(orc_scene* s, int x, int y, int rng_mode, uint64_t seed, double eps, double out[6]) {
    try {
        Scene& sc = *s->scene;
        sc.beginFrame();
        float dummy[3];
        Renderer r(sc, dummy);
        r.rngMode = rng_mode;
        r.seed = seed;
        Diag& dg = tl_diag();
        dg = Diag();
        dg.on = true;
        Color base = r.renderPixelShifted(x, y, 0.0, 0.0);
        dg.on = false;
        out[0] = dg.max_dist;
        out[1] = dg.min_gap;
        double worst = 0;
        for (int k = 0; k < 4; k++) {
            Color c = r.renderPixelShifted(x, y, (k & 1) ? eps : -eps, (k & 2) ? eps : -eps);
            worst = std::max(worst, (double)std::fabs(raw(c.r) - raw(base.r)));
            worst = std::max(worst, (double)std::fabs(raw(c.g) - raw(base.g)));
            worst = std::max(worst, (double)std::fabs(raw(c.b) - raw(base.b)));
        }
        out[2] = worst;
        out[3] = raw(base.r); out[4] = raw(base.g); out[5] = raw(base.b);
        return 0;
    } catch (const std::exception& e) {
        tl_diag().on = false;
        g_err = e.what();
        return -1;
    }
}

// 8-bit packing of a float frame through the reference LUT (color.d:154-162,194-229).
void orc_pack_rgb32(const float* rgb, size_t npx, uint32_t* out) {
    static const SrgbLut lut;
    for (size_t i = 0; i < npx; i++) out[i] = lut.toRGB32(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
}
void orc_srgb_lut(uint8_t out[4097]) {
    static const SrgbLut lut;
    memcpy(out, lut.t, 4097);
}

// ---- micro-KAT hooks (tests/test_oracle_kat.py) -----------------------------------------------
// type: 0 plane(y) 1 sphere(cx,cy,cz,R) 2 cube(cx,cy,cz,side).  out: hit, dist, p[3], n[3], u, v
int orc_kat_intersect(int type, const double params[4], const double o[3], const double d[3], double max_dist,
                      double out[10]) {
    std::unique_ptr<Geometry> g;
    if (type == 0) { auto p = std::make_unique<Plane>(); p->y = mk_real(params[0]); g = std::move(p); }
    else if (type == 1) {
        auto s = std::make_unique<Sphere>();
        s->center = Vec3(mk_real(params[0]), mk_real(params[1]), mk_real(params[2]));
        s->R = mk_real(params[3]);
        g = std::move(s);
    } else if (type == 2) {
        auto c = std::make_unique<Cube>();
        c->center = Vec3(mk_real(params[0]), mk_real(params[1]), mk_real(params[2]));
        c->side = mk_real(params[3]);
        g = std::move(c);
    } else return -1;
    Ray r;
    r.orig = Vec3(mk_real(o[0]), mk_real(o[1]), mk_real(o[2]));
    r.dir = Vec3(mk_real(d[0]), mk_real(d[1]), mk_real(d[2]));
    IntersectionData data;
    data.dist = mk_real(max_dist);
    bool hit = g->intersect(r, data);
    out[0] = hit;
    out[1] = raw(data.dist);
    out[2] = raw(data.p.x); out[3] = raw(data.p.y); out[4] = raw(data.p.z);
    out[5] = raw(data.normal.x); out[6] = raw(data.normal.y); out[7] = raw(data.normal.z);
    out[8] = raw(data.u); out[9] = raw(data.v);
    return 0;
}

void orc_kat_checker(double size, double u, double v, const float c1[3], const float c2[3], float out[3]) {
    Checker c;
    c.size = mk_real(size);
    c.color1 = Color::fromFloats(c1[0], c1[1], c1[2]);
    c.color2 = Color::fromFloats(c2[0], c2[1], c2[2]);
    Ray r;
    Vec3 n;
    Color k = c.getTexColor(r, mk_real(u), mk_real(v), n);
    out[0] = raw(k.r); out[1] = raw(k.g); out[2] = raw(k.b);
}

// shell sort of `n` distances; returns the permutation applied (util/array.d:95-111)
void orc_kat_shell_sort(const double* dist, int n, int* perm) {
    std::vector<IntersectionData> v((size_t)n);
    std::vector<Plane> tags((size_t)n);
    for (int i = 0; i < n; i++) { v[i].dist = mk_real(dist[i]); v[i].g = &tags[i]; }
    shell_sort(v);
    for (int i = 0; i < n; i++) perm[i] = (int)(static_cast<const Plane*>(v[i].g) - tags.data());
}

// BMP decode of an in-memory file -> packed 0xAARRGGBB (bmp.d KATs :446-611)
int orc_kat_decode_bmp(const uint8_t* bytes, size_t n, uint32_t* w, uint32_t* h, uint32_t* out, size_t out_cap) {
    try {
        std::vector<uint8_t> f(bytes, bytes + n);
        size_t W, H;
        std::vector<uint32_t> px;
        decode_bmp(f, W, H, px);
        *w = (uint32_t)W; *h = (uint32_t)H;
        if (px.size() > out_cap) { g_err = "output too small"; return -1; }
        memcpy(out, px.data(), px.size() * 4);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// texels of bitmap texture #idx after load-time gamma (float RGB); returns w,h
int orc_texture_texels(const orc_scene* s, int idx, uint32_t* w, uint32_t* h, float* out, size_t out_cap_floats) {
    if (idx < 0 || (size_t)idx >= s->scene->textures.size()) return -1;
    auto* b = dynamic_cast<const BitmapTexture*>(s->scene->textures[idx].get());
    if (!b) return -2;
    *w = (uint32_t)b->bmp.width; *h = (uint32_t)b->bmp.height;
    if (out) {
        if (b->bmp.px.size() > out_cap_floats) return -3;
        memcpy(out, b->bmp.px.data(), b->bmp.px.size() * sizeof(float));
    }
    return 0;
}

uint32_t orc_rng_u31(uint64_t seed, uint32_t px, uint32_t py, uint32_t tap, uint32_t sample, uint32_t draw) {
    return rng_u31(seed, px, py, tap, sample, draw);
}

int orc_counts_flops(void) {
#ifdef ORC_COUNT_FLOPS
    return 1;
#else
    return 0;
#endif
}

}  // extern "C"
