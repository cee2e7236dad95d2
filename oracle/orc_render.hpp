// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_scalar.hpp).
//
// Frame driver restated from /root/reference/source/rt/renderer.d:
//   renderRT :83-189 (prepass, 1-spp pass, AA pass over ALL pixels — the needsAA mask is dead),
//   getBucketsList :194-213, drawRect :215-220, renderPixelNoAA :223-228, renderPixelAA :233-251,
//   renderSample :254-268, renderSampleDof :270-287, renderSampleDefault :303-313,
//   trace :325-358, raytrace_impl :361-376; environment.d:7-10 (miss = black).
//   renderSampleGI :289-301, pathtrace :320-323, pathtrace_impl :378-463 (GI branch).
// Worker scheduling mirrors std.parallelism's dynamic hand-out of
// buckets (renderer.d:133-136) with a shared atomic bucket counter.
#pragma once
#include <atomic>
#include <exception>
#include <mutex>
#include <thread>

#include "orc_scene.hpp"

namespace orc {

struct Box2i {
    int x0, y0, x1, y1;
};

struct Renderer {
    const Scene& scene;
    float* out;  // W*H*3, row-major, top row first (imageio/image.d:45-54)
    uint32_t W, H;
    int rngMode = 1;
    uint64_t seed = 0;

    Renderer(const Scene& s, float* o) : scene(s), out(o), W(s.settings.frameWidth), H(s.settings.frameHeight) {}

    std::vector<Box2i> getBucketsList() const {  // renderer.d:194-213
        const int BS = (int)scene.settings.bucketSize;
        int w = (int)W, h = (int)H;
        int BW = (w - 1) / BS + 1, BH = (h - 1) / BS + 1;
        std::vector<Box2i> res;
        for (int y = 0; y < BH; y++) {
            if (y % 2 == 0)
                for (int x = 0; x < BW; x++) res.push_back({x * BS, y * BS, (x + 1) * BS, (y + 1) * BS});
            else
                for (int x = BW - 1; x >= 0; x--) res.push_back({x * BS, y * BS, (x + 1) * BS, (y + 1) * BS});
        }
        for (auto& b : res) {  // imported_types.d:31-35 clip
            b.x1 = std::min(b.x1, w);
            b.y1 = std::min(b.y1, h);
        }
        return res;
    }

    Color getPx(int x, int y) const {
        const float* q = &out[((size_t)W * y + x) * 3];
        return Color::fromFloats(q[0], q[1], q[2]);
    }
    void setPx(int x, int y, const Color& c) {
        float* q = &out[((size_t)W * y + x) * 3];
        q[0] = raw(c.r); q[1] = raw(c.g); q[2] = raw(c.b);
    }

    Color trace(const Ray& ray) const {  // renderer.d:325-376
        if ((uint32_t)ray.depth > scene.settings.maxTraceDepth) return Color::fromFloats(0, 0, 0);
        IntersectionData data;
        data.dist = mk_real(1e99);
        const Node* closestNode = nullptr;
        for (auto& node : scene.nodes)
            if (node->intersect(ray, data)) closestNode = node.get();
        if (tl_diag().on && closestNode) {   // conditioning probe (orc_pixel_diag): hit distance, nearest competing node
            Diag& dg = tl_diag();
            const double d1 = raw(data.dist);
            if (d1 > dg.max_dist) dg.max_dist = d1;
            for (auto& node : scene.nodes) {
                if (node.get() == closestNode) continue;
                IntersectionData probe;
                probe.dist = mk_real(1e99);
                if (node->intersect(ray, probe)) dg.gap(std::fabs(raw(probe.dist) - d1) / std::max(1.0, std::fabs(d1)));
            }
        }
        // lights are never intersectable (light.d:67-70) -> hitLight stays false
        if (!closestNode) return scene.environment.getEnvironment(ray.dir);  // renderer.d:366-368; environment.d:7-10 (black) or the cubemap extension
        // bumpmap.modifyNormal is a no-op (texture.d:10-12)
        return closestNode->shader->shade(ray, data);
    }

    // renderer.d:320-358 with TraceType.Path.  `pathtrace` drops its pathMultiplier argument and `trace` restarts
    // pathtrace_impl from Color(1, 1, 1) (:322, :356) — reproduced: the multiplier never accumulates.
    Color pathtrace(const Ray& ray) const {
        if ((uint32_t)ray.depth > scene.settings.maxTraceDepth) return Color::fromFloats(0, 0, 0);
        IntersectionData data;
        data.dist = mk_real(1e99);
        const Node* closestNode = nullptr;
        for (auto& node : scene.nodes)
            if (node->intersect(ray, data)) closestNode = node.get();
        // PointLight.intersect is false (light.d:67-70): hitLight never set, the :380-393 branch is dead
        const Color pathMultiplier = Color::fromFloats(1, 1, 1);
        if (!closestNode) return scene.environment.getEnvironment(ray.dir) * pathMultiplier;  // :396-397, environment.d:7-10
        Color resultDirect = Color::fromFloats(0, 0, 0);
        if (!scene.lights.empty()) {  // :404-445
            const size_t lightIndex = uniformIndex(scene.lights.size());
            const PointLight& light = *scene.lights[lightIndex];
            const size_t lightSampleIdx = uniformIndex(light.getNumSamples());
            Vec3 pointOnLight;
            Color lightColor;
            light.getNthSample(lightSampleIdx, data.p, pointOnLight, lightColor);
            if (raw(lightColor.intensity()) > 0 && scene.testVisibility(data.p + data.normal * mk_real(1e-6), pointOnLight)) {
                Ray w_out;
                w_out.orig = data.p + data.normal * mk_real(1e-6);
                w_out.dir = pointOnLight - w_out.orig;
                normalize(w_out.dir);
                const float solidAngle = light.solidAngle(w_out.orig);
                Color brdfAtPoint = closestNode->shader->eval(data, ray, w_out);
                lightColor = light.color() * mk_colf(solidAngle) / mk_colf((float)(2 * PI_L));
                const float pdfChooseLight = 1.0f / (float)scene.lights.size();
                const float pdfInLight = (float)(1 / (2 * PI_L));
                const float pdf = pdfChooseLight * pdfInLight;
                if (raw(brdfAtPoint.intensity()) > 0) resultDirect = lightColor * pathMultiplier * brdfAtPoint / mk_colf(pdf);
            }
        }
        Ray w_out;
        Color brdfEval;
        float pdf;
        closestNode->shader->spawnRay(data, ray, w_out, brdfEval, pdf);  // :452
        if (pdf < 0) return Color::fromFloats(1, 0, 0);
        if (pdf == 0) return Color::fromFloats(0, 0, 0);
        tl_stats().bounce++;
        Color resultGi = pathtrace(w_out);  // :458 (the multiplier argument is dropped by pathtrace)
        return resultDirect + resultGi;
    }

    Color renderSample(real x, real y, int dx, int dy, uint32_t tap) const {  // renderer.d:254-313
        RngState& rs = tl_rng();
        rs.tap = tap;
        const bool stereo = scene.camera.stereoSeparation != 0;
        if (scene.camera.dof) {
            Color average = Color::fromFloats(0, 0, 0);
            for (size_t i = 0; i < scene.camera.numSamples; i++) {
                rs.sample = (uint32_t)i;
                rs.draw = 0;
                if (!stereo) {
                    real jx = x + uniform01() * mk_real((double)dx);
                    real jy = y + uniform01() * mk_real((double)dy);
                    tl_stats().primary++;
                    average += trace(scene.camera.getScreenRay(jx, jy));
                } else {  // renderer.d:280-283: each eye draws its own jitter and lens sample
                    real lx = x + uniform01() * mk_real((double)dx);
                    real ly = y + uniform01() * mk_real((double)dy);
                    Color left = trace(scene.camera.getScreenRay(lx, ly, -1));
                    real rx = x + uniform01() * mk_real((double)dx);
                    real ry = y + uniform01() * mk_real((double)dy);
                    Color right = trace(scene.camera.getScreenRay(rx, ry, +1));
                    tl_stats().primary += 2;
                    average += combineStereo(left, right);
                }
            }
            return average / mk_colf((float)scene.camera.numSamples);
        }
        if (scene.settings.GIEnabled) {  // renderer.d:260-263,289-301 (after the dof test: DOF wins)
            Color average = Color::fromFloats(0, 0, 0);
            for (uint32_t i = 0; i < scene.settings.pathsPerPixel; i++) {
                rs.sample = i;
                rs.draw = 0;
                real jx = x + uniform01() * mk_real((double)dx);
                real jy = y + uniform01() * mk_real((double)dy);
                tl_stats().primary++;
                average += pathtrace(scene.camera.getScreenRay(jx, jy));
            }
            return average / mk_colf((float)scene.settings.pathsPerPixel);
        }
        if (stereo) {  // renderer.d:307-312
            tl_stats().primary += 2;
            return combineStereo(trace(scene.camera.getScreenRay(x, y, -1)), trace(scene.camera.getScreenRay(x, y, +1)));
        }
        tl_stats().primary++;
        return trace(scene.camera.getScreenRay(x, y));
    }

    Color renderPixelNoAA(int x, int y, int dx = 1, int dy = 1, uint32_t tap = 0) {  // renderer.d:223-228
        RngState& rs = tl_rng();
        rs.mode = rngMode; rs.seed = seed; rs.px = (uint32_t)x; rs.py = (uint32_t)y;
        Color result = renderSample(mk_real((double)x), mk_real((double)y), dx, dy, tap);
        setPx(x, y, result);
        return result;
    }

    Color renderPixelAA(int x, int y) {  // renderer.d:233-251
        static const double kernel[5][2] = {{0.0, 0.0}, {0.3, 0.3}, {0.6, 0.0}, {0.0, 0.6}, {0.6, 0.6}};
        RngState& rs = tl_rng();
        rs.mode = rngMode; rs.seed = seed; rs.px = (uint32_t)x; rs.py = (uint32_t)y;
        Color accum = getPx(x, y);
        for (int sample = 1; sample < 5; sample++)
            accum += renderSample(mk_real((double)x) + mk_real(kernel[sample][0]), mk_real((double)y) + mk_real(kernel[sample][1]), 1, 1,
                                  (uint32_t)sample);
        setPx(x, y, accum / mk_colf(5.f));
        return getPx(x, y);
    }

    // One whole pixel (corner sample + the AA taps) with every sample position shifted by (ox, oy) pixels; writes nothing.
    // Used by the conditioning probe: a pixel whose colour moves under a 1e-7 pixel shift sits on a discontinuity of the
    // reference's own image.
    Color renderPixelShifted(int x, int y, double ox, double oy) const {
        static const double kernel[5][2] = {{0.0, 0.0}, {0.3, 0.3}, {0.6, 0.0}, {0.0, 0.6}, {0.6, 0.6}};
        RngState& rs = tl_rng();
        rs.mode = rngMode; rs.seed = seed; rs.px = (uint32_t)x; rs.py = (uint32_t)y;
        const int taps = scene.settings.AAEnabled ? 5 : 1;
        Color accum = Color::fromFloats(0, 0, 0);
        for (int t = 0; t < taps; t++) {
            Color c = renderSample(mk_real((double)x + ox) + mk_real(kernel[t][0]), mk_real((double)y + oy) + mk_real(kernel[t][1]), 1, 1, (uint32_t)t);
            if (t == 0) accum = c;
            else accum += c;
        }
        return taps == 5 ? accum / mk_colf(5.f) : accum;
    }

    template <class F>
    void parallelBuckets(const std::vector<Box2i>& buckets, unsigned nthreads, Stats& total, F&& body) {
        std::atomic<size_t> next{0};
        std::mutex m;
        std::exception_ptr failure;
        auto worker = [&]() {
            tl_stats() = Stats();
#ifdef ORC_COUNT_FLOPS
            FlopCounter::tl() = 0;
#endif
            try {
                for (;;) {
                    size_t i = next.fetch_add(1);
                    if (i >= buckets.size()) break;
                    const Box2i& b = buckets[i];
                    for (int y = b.y0; y < b.y1; y++)
                        for (int x = b.x0; x < b.x1; x++) body(x, y);
                }
            } catch (...) {   // ReferenceHalts (GI on a Phong surface): reported by the calling thread
                std::lock_guard<std::mutex> g(m);
                if (!failure) failure = std::current_exception();
                next.store(buckets.size());
            }
            std::lock_guard<std::mutex> g(m);
            total.primary += tl_stats().primary;
            total.shadow += tl_stats().shadow;
            total.bounce += tl_stats().bounce;
            total.csg_max_crossings = std::max(total.csg_max_crossings, tl_stats().csg_max_crossings);
#ifdef ORC_COUNT_FLOPS
            total.flops += FlopCounter::tl();
#endif
        };
        if (nthreads <= 1) {
            worker();
        } else {
            std::vector<std::thread> th;
            for (unsigned t = 0; t < nthreads; t++) th.emplace_back(worker);
            for (auto& t : th) t.join();
        }
        if (failure) std::rethrow_exception(failure);
    }

    // Returns the ray/flop counts of the passes that reach the final image (pass 2 + pass 3);
    // prepass work is timed (it is part of the reference's frame) but counted separately.
    void renderRT(unsigned nthreads, Stats& finalStats, Stats& prepassStats) {  // renderer.d:83-189
        std::vector<Box2i> buckets = getBucketsList();
        if (scene.settings.prepassEnabled) {  // :110-127, serial on the calling thread
            tl_stats() = Stats();
#ifdef ORC_COUNT_FLOPS
            FlopCounter::tl() = 0;
#endif
            for (auto& r : buckets) {
                int rw = r.x1 - r.x0, rh = r.y1 - r.y0;
                for (int dy = 0; dy < rh; dy += 16) {
                    int ey = std::min(rh, dy + 16);
                    for (int dx = 0; dx < rw; dx += 16) {
                        int ex = std::min(rw, dx + 16);
                        Color c = renderPixelNoAA(r.x0 + dx, r.y0 + dy, ex - dx, ey - dy);
                        for (int yy = r.y0 + dy; yy < r.y0 + ey; yy++)
                            for (int xx = r.x0 + dx; xx < r.x0 + ex; xx++) setPx(xx, yy, c);
                    }
                }
            }
            prepassStats = tl_stats();
#ifdef ORC_COUNT_FLOPS
            prepassStats.flops = FlopCounter::tl();
#endif
        }
        if (scene.settings.prepassOnly) return;
        parallelBuckets(buckets, nthreads, finalStats, [&](int x, int y) { renderPixelNoAA(x, y); });  // :133-142
        if (!scene.settings.AAEnabled) return;
        // :150-178 computes a needsAA mask that nothing reads afterwards; kept so the CPU baseline
        // pays for it like the reference does.
        std::vector<uint8_t> needsAA((size_t)W * H, 0);
        parallelBuckets(buckets, nthreads, finalStats, [&](int x, int y) {
            int xs[5] = {x, x > 0 ? x - 1 : x, x + 1 < (int)W ? x + 1 : x, x, x};
            int ys[5] = {y, y, y, y > 0 ? y - 1 : y, y + 1 < (int)H ? y + 1 : y};
            float n[5][3], avg[3] = {0, 0, 0};
            for (int i = 0; i < 5; i++) {
                const float* q = &out[((size_t)W * ys[i] + xs[i]) * 3];
                for (int c = 0; c < 3; c++) { n[i][c] = q[c]; avg[c] += q[c]; }
            }
            const float rdiv = 1.0f / 5.0f;  // color.d:109-118 `/=` multiplies by the reciprocal
            for (int c = 0; c < 3; c++) avg[c] *= rdiv;
            for (int i = 0; i < 5; i++)
                if (std::fabs(n[i][0] - avg[0]) > 0.1f || std::fabs(n[i][1] - avg[1]) > 0.1f || std::fabs(n[i][2] - avg[2]) > 0.1f) {
                    needsAA[(size_t)W * y + x] = 1;
                    break;
                }
        });
        parallelBuckets(buckets, nthreads, finalStats, [&](int x, int y) { renderPixelAA(x, y); });  // :183-186
    }
};

}  // namespace orc
