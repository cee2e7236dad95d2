// Hand-written sm_100a render kernel: one thread per pixel, all AA taps / DOF samples of a pixel
// in the same thread, scene in constant memory, warp-coherent 8x4 pixel patches, float4-vectorised
// framebuffer stores staged through shared memory.
//
// What each device function replaces in the reference (paths relative to /root/reference/source):
//   gen_ray            rt/camera.d:123-174 getScreenRay (+ :258-269 unitDiscSample via sincos_rev, util/random.d:19-28)
//   isect_plane        rt/geometry.d:30-59 (isect_plane_u: the same test for the un-normalised camera rays of plane-only scenes)
//   isect_sphere       rt/geometry.d:92-125
//   isect_cube         rt/geometry.d:172-235
//   isect_csg          rt/geometry.d:271-332, :382-397; util/array.d:95-111 (shell sort) — closed form, primitive children
//   isect_geom_lit     the same walk replayed literally for CSG nested inside CSG
//   geom_inside        rt/geometry.d:25-28,127-130,165-170,334-337
//   node_hit           rt/node.d:23-49 + rt/transform.d:57-86 around the primitive / CSG test (node_exact: plane-only scene classes)
//   cull_sphere        (no counterpart: conservative FP32 bounding-sphere rejection per ray)
//   camera_mask / shadow_mask  (no counterpart: per-warp node masks by warp ballot — which nodes can ANY ray of the warp reach)
//   trace_warp         rt/renderer.d:325-376 trace + rt/shader.d:67-105,197-250 Lambert / Phong + rt/scene.d:62-78 testVisibility:
//                      camera ray and shadow rays share ONE warp-uniform node walk over the masks; shadow walks are left
//                      through __all_sync (trace / shade / occluded_planes: plane-only scenes, shadows settled by the sign of D.y)
//   sample_texture     rt/texture.d:36-54,77-86 (Procedure2's sines through sin_rev),116-126; bitmap_fetch rt/bitmap.d:48-63
//   env_lookup         rt/environment.d:7-10 (black) + the cubemap EXTENSION (no counterpart in the reference)
//   render_sample      rt/renderer.d:254-313 (renderSampleDefault / renderSampleDof, stereo via color.d:10-15)
//   render_frame_kernel rt/renderer.d:83-251 (renderRT: 1-spp + AA passes fused per pixel; prepassOnly preview;
//                      GI frames, renderer.d:289-301,378-463, are black by construction: fp.gi, see c2rt_api.cu fill_params)
//   render_pixel_kernel rt/renderer.d:46-57 (renderPixel)
//   pack_rgb32         rt/color.d:154-162,209-214
// Decisions and coordinates run in FP64, colours in FP32 (precision plan below).
// The kernel is instantiated per scene class (MODE_* in scene_dev.h): one plane + one light scenes (MODE_SOLO) address
// every scene constant statically and compile the texture / shader kind in; launch_frame at the end of the file dispatches.
#include <cuda_runtime.h>
#include <math_constants.h>

#include "scene_dev.h"

namespace c2rt {

__constant__ DevScene c_scene;
__constant__ double c_tap_x[5] = {0.0, 0.3, 0.6, 0.0, 0.6};  // renderer.d:235-242
__constant__ double c_tap_y[5] = {0.0, 0.3, 0.0, 0.6, 0.6};

// Scene records.  Scenes that fit the constant block (C2RT_MAX_*) are read from it: every record read in the warp-uniform node
// walk is a constant-cache broadcast.  Larger scenes (MODE_BIG) keep the same records in global memory (c_scene.g_*): the walk's
// reads are still one address per warp, served by L1.
template <bool BIG> __device__ __forceinline__ const DevNode& node_at(int i) {
    if constexpr (BIG) { __builtin_assume(__isGlobal(c_scene.g_nodes)); return c_scene.g_nodes[i]; } else return c_scene.nodes[i];
}
template <bool BIG> __device__ __forceinline__ const DevGeom& geom_at(int i) {
    if constexpr (BIG) { __builtin_assume(__isGlobal(c_scene.g_geoms)); return c_scene.g_geoms[i]; } else return c_scene.geoms[i];
}
template <bool BIG> __device__ __forceinline__ const DevShader& shader_at(int i) {
    if constexpr (BIG) { __builtin_assume(__isGlobal(c_scene.g_shaders)); return c_scene.g_shaders[i]; } else return c_scene.shaders[i];
}
template <bool BIG> __device__ __forceinline__ const DevTex& tex_at(int i) {
    if constexpr (BIG) { __builtin_assume(__isGlobal(c_scene.g_textures)); return c_scene.g_textures[i]; } else return c_scene.textures[i];
}
__host__ __device__ constexpr bool is_big(int mode) { return (mode & MODE_BIG) != 0; }

// Precision plan: FP64 carries everything a pixel DECISION or a texture coordinate
// depends on — ray direction, hit distances, hit points, plane/cube uv, checker cells, face-forward
// sign, CSG crossing order.  FP32 carries what is continuous and ends in an FP32 colour anyway —
// bounding-sphere culling (conservative margins), normals for lighting, light falloff, cosines, pow,
// sin after an FP64 range reduction, sphere uv.  Divisions and square roots in FP64 use the
// MUFU seed + Newton steps below instead of the IEEE library sequences (|rel err| < 4e-16).

// FP64 literals of the node walk as constant-bank operands: a DSETP / DFMA / DADD reads c[bank][imm] directly, whereas a literal
// whose low word is not zero is first assembled in a register pair by two moves at every use
__constant__ double c_lit[4] = {1e-9, 1e-6, 1e300, 1e99};
#ifdef C2RT_LITERAL_IMMEDIATES
#define K_1EM9 1e-9
#define K_1EM6 1e-6
#define K_1E300 1e300
#define K_1E99 1e99
#else
#define K_1EM9 c_lit[0]
#define K_1EM6 c_lit[1]
#define K_1E300 c_lit[2]
#define K_1E99 c_lit[3]
#endif

struct Ray {
    double ox, oy, oz, dx, dy, dz;   // d is unit (to ~1e-16) — except camera rays of the plane-only scene classes, see l2
    double l2;                       // plane-only scene classes: camera rays keep the UN-normalised direction, l2 = |d|^2 (gen_ray)
    float fox, foy, foz, fdx, fdy, fdz, olen;  // FP32 shadow of the ray for the conservative cull
};

// Closest-hit record.  `p` is in the node's local frame: world space for KIND_*_W nodes, object
// space for KIND_GENERIC.  Normal / uv are derived from (leaf, face, p) for the winning hit only
// (the reference fills them for every candidate: geometry.d:49-55,114-123,224-230).
struct HitRec {
    double dist;
    double px, py, pz;
    int node, leaf, face;
};

constexpr int FACE_FLIP = 8;  // CsgDiff normal flip (geometry.d:394-395)

struct Col {
    float r, g, b;
};
__device__ __forceinline__ Col mkcol(float r, float g, float b) { Col c; c.r = r; c.g = g; c.b = b; return c; }

__device__ __forceinline__ double dot3(double ax, double ay, double az, double bx, double by, double bz) {
    return fma(az, bz, fma(ay, by, ax * bx));
}
__device__ __forceinline__ float dot3f(float ax, float ay, float az, float bx, float by, float bz) {
    return fmaf(az, bz, fmaf(ay, by, ax * bx));
}
// sqrt(a) from the SFU, nudged up: an upper bound (to ~1e-6) for quantities that only feed conservative margins / radii
__device__ __forceinline__ float sqrt_up(float a) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r * 1.000002f;
}
// 1/sqrt(a) in FP32: MUFU.RSQ as rsqrtf issues it, without rsqrtf's rescaling of denormal inputs (three more instructions per
// call; the arguments here are squared lengths of non-degenerate vectors).  Identical result for every normal input.
__device__ __forceinline__ float rsqrt_fast(float a) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}
// 1/a and 1/sqrt(a) in FP64: hardware seed (MUFU.RCP64H / MUFU.RSQ64H, ~20 bits) + two Newton steps
__device__ __forceinline__ double rcp64(double a) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double e = fma(-a, y, 1.0);
    y = fma(y, e, y);
    e = fma(-a, y, 1.0);
    y = fma(y, e, y);
    return y;
}
__device__ __forceinline__ double rsqrt64(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double h = 0.5 * y;
    double e = fma(-a * y, y, 1.0);
    y = fma(h, e, y);
    h = 0.5 * y;
    e = fma(-a * y, y, 1.0);
    y = fma(h, e, y);
    return y;
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void normalize3(double& x, double& y, double& z) {
    double inv = rsqrt64(dot3(x, y, z, x, y, z));
    x *= inv; y *= inv; z *= inv;
}
// row vector x row-major 3x3 (imported_types.d:13-20)
__device__ __forceinline__ void mulvm(const double* m, double x, double y, double z, double& rx, double& ry, double& rz) {
    rx = x * m[0] + y * m[3] + z * m[6];
    ry = x * m[1] + y * m[4] + z * m[7];
    rz = x * m[2] + y * m[5] + z * m[8];
}
// double -> float that the register allocator may not rematerialise (F2F runs on the quarter-rate XU pipe;
// under a register cap ptxas otherwise re-converts at every use inside the node loop)
__device__ __forceinline__ float cvt_keep(double x) {
#ifdef C2RT_NO_KEEP
    return (float)x;
#else
    float f;
    asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f) : "d"(x));
    return f;
#endif
}
// FP32 shadow of a camera ray for the conservative cull.  `r.d` is still UN-normalised here: the FP32 direction is
// normalised in FP32 from it (error ~3 ulp, inside the cull's margins) instead of being a conversion of the FP64 unit
// vector — ptxas rematerialises such a conversion (F2F, quarter-rate XU pipe) in every iteration of the node loops.
template <int MODE>
__device__ __forceinline__ void set_shadow(const FrameParams& fp, Ray& r) {
    if (MODE & MODE_SAMPLING) {   // DOF / stereo move the origin per ray
        r.fox = cvt_keep(r.ox); r.foy = cvt_keep(r.oy); r.foz = cvt_keep(r.oz);
        r.olen = sqrtf(dot3f(r.fox, r.foy, r.foz, r.fox, r.foy, r.foz));
    } else {                      // the camera position: converted once per frame on the host
        r.fox = fp.posf[0]; r.foy = fp.posf[1]; r.foz = fp.posf[2];
        r.olen = fp.posf_len;
    }
    const float vx = (float)r.dx, vy = (float)r.dy, vz = (float)r.dz;
    const float inv = rsqrt_fast(dot3f(vx, vy, vz, vx, vy, vz));
    r.fdx = vx * inv; r.fdy = vy * inv; r.fdz = vz * inv;
}

// ---------------------------------------------------------------- pinned RNG (c2rt.h c2rt_rng_u31)
__device__ __forceinline__ uint32_t rng_u31(unsigned long long seed, uint32_t px, uint32_t py, uint32_t tap, uint32_t sample,
                                            uint32_t draw) {
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
__device__ __forceinline__ double uniform01(const FrameParams& fp, uint32_t px, uint32_t py, uint32_t tap, uint32_t sample,
                                            uint32_t& draw) {
    // u31 -> double through the 2^52 mantissa trick (one DADD; I2F.F64 would go through the quarter-rate XU pipe)
    const double r = __hiloint2double(0x43300000, (int)rng_u31(fp.seed, px, py, tap, sample, draw++)) - 4503599627370496.0;
    return r * (1.0 / 2147483647.0);
}

// sin(2 pi u), cos(2 pi u) for u in [0, 1] in FP64 (|err| < 4e-16): u = k / 64 + r with |r| <= 1 / 128 by the 2^52 rounding
// trick, the rotation by k / 64 of a revolution from a 64-entry table (correctly rounded, global memory: the index differs per
// lane, the 1 KB stays in L1), and short Taylor kernels on |t| = |2 pi r| <= 0.0491 (first neglected terms: t^11 / 11! and
// t^10 / 10!, < 1e-19).  Replaces sincos(u * 2 pi) of camera.d:260-263 — whose library form spends most of its ~75 instructions
// on argument ranges a unit-interval input cannot reach — with 19 FP64 instructions (round 1's quadrant form: 24 + 20 selects).
__device__ const double2 c_rot64[64] = {
    {1.0, 0.0}, {0.9951847266721969, 0.0980171403295606},
    {0.9807852804032304, 0.19509032201612828}, {0.9569403357322088, 0.2902846772544624},
    {0.9238795325112867, 0.3826834323650898}, {0.881921264348355, 0.47139673682599764},
    {0.8314696123025452, 0.5555702330196022}, {0.773010453362737, 0.6343932841636455},
    {0.7071067811865476, 0.7071067811865476}, {0.6343932841636455, 0.773010453362737},
    {0.5555702330196022, 0.8314696123025452}, {0.47139673682599764, 0.881921264348355},
    {0.3826834323650898, 0.9238795325112867}, {0.2902846772544624, 0.9569403357322088},
    {0.19509032201612828, 0.9807852804032304}, {0.0980171403295606, 0.9951847266721969},
    {0.0, 1.0}, {-0.0980171403295606, 0.9951847266721969},
    {-0.19509032201612828, 0.9807852804032304}, {-0.2902846772544624, 0.9569403357322088},
    {-0.3826834323650898, 0.9238795325112867}, {-0.47139673682599764, 0.881921264348355},
    {-0.5555702330196022, 0.8314696123025452}, {-0.6343932841636455, 0.773010453362737},
    {-0.7071067811865476, 0.7071067811865476}, {-0.773010453362737, 0.6343932841636455},
    {-0.8314696123025452, 0.5555702330196022}, {-0.881921264348355, 0.47139673682599764},
    {-0.9238795325112867, 0.3826834323650898}, {-0.9569403357322088, 0.2902846772544624},
    {-0.9807852804032304, 0.19509032201612828}, {-0.9951847266721969, 0.0980171403295606},
    {-1.0, 0.0}, {-0.9951847266721969, -0.0980171403295606},
    {-0.9807852804032304, -0.19509032201612828}, {-0.9569403357322088, -0.2902846772544624},
    {-0.9238795325112867, -0.3826834323650898}, {-0.881921264348355, -0.47139673682599764},
    {-0.8314696123025452, -0.5555702330196022}, {-0.773010453362737, -0.6343932841636455},
    {-0.7071067811865476, -0.7071067811865476}, {-0.6343932841636455, -0.773010453362737},
    {-0.5555702330196022, -0.8314696123025452}, {-0.47139673682599764, -0.881921264348355},
    {-0.3826834323650898, -0.9238795325112867}, {-0.2902846772544624, -0.9569403357322088},
    {-0.19509032201612828, -0.9807852804032304}, {-0.0980171403295606, -0.9951847266721969},
    {0.0, -1.0}, {0.0980171403295606, -0.9951847266721969},
    {0.19509032201612828, -0.9807852804032304}, {0.2902846772544624, -0.9569403357322088},
    {0.3826834323650898, -0.9238795325112867}, {0.47139673682599764, -0.881921264348355},
    {0.5555702330196022, -0.8314696123025452}, {0.6343932841636455, -0.773010453362737},
    {0.7071067811865476, -0.7071067811865476}, {0.773010453362737, -0.6343932841636455},
    {0.8314696123025452, -0.5555702330196022}, {0.881921264348355, -0.47139673682599764},
    {0.9238795325112867, -0.3826834323650898}, {0.9569403357322088, -0.2902846772544624},
    {0.9807852804032304, -0.19509032201612828}, {0.9951847266721969, -0.0980171403295606},
};
// (coefficients in constant memory: a DFMA takes them as c[bank][imm] operands; FP64 literals would each cost two moves)
__constant__ double c_sincos[12] = {
    2.75573192239858906525573e-6, -1.98412698412698412698413e-4, 8.33333333333333333333333e-3, -1.66666666666666666666667e-1,   // 1/9! -1/7! 1/5! -1/3!
    2.48015873015873015873016e-5, -1.38888888888888888888889e-3, 4.16666666666666666666667e-2, -0.5,                            // 1/8! -1/6! 1/4! -1/2!
    6.283185307179586476925, 6755399441055744.0, -0.015625, 64.0};
__device__ __forceinline__ void sincos_rev(double u, double& s, double& c) {
    const double MAGIC = c_sincos[9];                   // 1.5 * 2^52
    const double qm = fma(u, c_sincos[11], MAGIC);      // nearest integer to 64 u in the low mantissa bits
    const double2 R = __ldg(&c_rot64[__double2loint(qm) & 63]);   // (cos, sin) of k / 64 revolutions; k = 64 is k = 0
    const double r = fma(qm - MAGIC, c_sincos[10], u);  // exact: u - k / 64 in [-1/128, 1/128]
    const double t = r * c_sincos[8];
    const double z = t * t;
    double ps = c_sincos[0];
    ps = fma(ps, z, c_sincos[1]);
    ps = fma(ps, z, c_sincos[2]);
    ps = fma(ps, z, c_sincos[3]);
    const double st = fma(t * z, ps, t);
    double pc = c_sincos[4];
    pc = fma(pc, z, c_sincos[5]);
    pc = fma(pc, z, c_sincos[6]);
    pc = fma(pc, z, c_sincos[7]);
    const double ct = fma(pc, z, 1.0);
    s = fma(R.x, st, R.y * ct);
    c = fma(-R.y, st, R.x * ct);
}

// ---------------------------------------------------------------- camera
// camera.d:123-174.  fp.up_left is stored relative to the camera position, fp.inv_w/inv_h are the
// reciprocals of the camera frame size (x / W -> x * (1/W): one rounding apart).
// Un-normalised direction through screen position (x, y): (upLeft - pos) + du * x/W + dv * y/H
__device__ __forceinline__ void screen_dir(const FrameParams& fp, double x, double y, double& vx, double& vy, double& vz) {
    double sx = x * fp.inv_w, sy = y * fp.inv_h;
    vx = fma(fp.dv[0], sy, fma(fp.du[0], sx, fp.ul_rel[0]));
    vy = fma(fp.dv[1], sy, fma(fp.du[1], sx, fp.ul_rel[1]));
    vz = fma(fp.dv[2], sy, fma(fp.du[2], sx, fp.ul_rel[2]));
}

// Completes a camera ray from its un-normalised direction (DOF lens sampling included)
// `eye`: 0 = Stereo3DOffset.None, -1 = Left, +1 = Right (camera.d:149-152,168-170)
template <int MODE>
__device__ __forceinline__ void gen_ray(const FrameParams& fp, double vx, double vy, double vz, uint32_t px, uint32_t py, uint32_t tap,
                                        uint32_t sample, uint32_t& draw, int eye, Ray& r) {
    // Plane-only scene classes keep camera rays UN-normalised: a plane hit o + d t and the closest-hit order do not
    // depend on |d| (all planes see the same parametrisation), the 1e-9 grazing thresholds of geometry.d:35-36 are
    // applied to d.y^2 / |d|^2 (isect_plane_u), and the DOF focal point o + d (f / (d . front)) is the same point for
    // any positive rescale of d.  This drops both FP64 normalisations (camera.d:144,172) from the ray.
    constexpr bool UNNORM = plane_only(MODE);
    r.dx = vx; r.dy = vy; r.dz = vz;
    r.ox = fp.pos[0]; r.oy = fp.pos[1]; r.oz = fp.pos[2];
    const double sep = eye > 0 ? fp.stereo_sep : -fp.stereo_sep;
    const bool stereo = (MODE & MODE_SAMPLING) && fp.stereo_sep != 0 && eye != 0;   // (warp-uniform first term: a branch, not predication)
    if (stereo) { r.ox += fp.right_dir[0] * sep; r.oy += fp.right_dir[1] * sep; r.oz += fp.right_dir[2] * sep; }
    if ((MODE & MODE_BOUNDED) && !((MODE & MODE_SAMPLING) && fp.dof)) set_shadow<MODE>(fp, r);   // (a DOF ray gets its final origin and direction below)
    if (!UNNORM) normalize3(r.dx, r.dy, r.dz);
    if ((MODE & MODE_SAMPLING) && fp.dof) {
        double cosTheta = dot3(r.dx, r.dy, r.dz, fp.front_dir[0], fp.front_dir[1], fp.front_dir[2]);
        double M = fp.focal_plane_dist * rcp64(cosTheta);
        double Tx = r.ox + r.dx * M, Ty = r.oy + r.dy * M, Tz = r.oz + r.dz * M;
        double u1 = uniform01(fp, px, py, tap, sample, draw);   // angle = u1 * 2 pi (camera.d:260)
        double u2 = uniform01(fp, px, py, tap, sample, draw);
        double rad = u2 > 0 ? u2 * rsqrt64(u2) : 0.0;
        double sa, ca;
        // lens position: kept in FP64.  It moves the ray origin by up to discMultiplier and so every texture coordinate of the
        // sample; an FP32 lens sample (measured: C3 14.9 -> 12.4 ms) shifts them by ~1e-7 of a texel, which is invisible
        // except where the reference's own image is discontinuous — bitmap.d:48-63 returns red when float(u) rounds up to 1.0 —
        // and one such sample in 8 M moved a zaphod pixel by 3e-3
        sincos_rev(u1, sa, ca);
        double ddx = sa * rad * fp.disc_multiplier;
        double ddy = ca * rad * fp.disc_multiplier;
        r.ox = fp.pos[0] + ddx * fp.right_dir[0] + ddy * fp.up_dir[0];
        r.oy = fp.pos[1] + ddx * fp.right_dir[1] + ddy * fp.up_dir[1];
        r.oz = fp.pos[2] + ddx * fp.right_dir[2] + ddy * fp.up_dir[2];
        if (stereo) { r.ox += fp.right_dir[0] * sep; r.oy += fp.right_dir[1] * sep; r.oz += fp.right_dir[2] * sep; }
        r.dx = Tx - r.ox; r.dy = Ty - r.oy; r.dz = Tz - r.oz;
        if (MODE & MODE_BOUNDED) set_shadow<MODE>(fp, r);
        if (!UNNORM) normalize3(r.dx, r.dy, r.dz);
    }
    // |d|^2 of an un-normalised camera ray: read by the general plane test and by Phong lobes; the one-plane kernels without
    // sampling form it only where they need it (isect_plane_solo)
    if (UNNORM && (!(MODE & MODE_SOLO) || (MODE & (MODE_SAMPLING | MODE_PHONG)))) r.l2 = dot3(r.dx, r.dy, r.dz, r.dx, r.dy, r.dz);
}

// ---------------------------------------------------------------- primitives
// Each returns true iff it found a hit with t <= dist, then updates dist and the hit point.
// (o, d) is the ray in the primitive's frame, p[] the primitive's parameters in that frame.
__device__ __forceinline__ bool isect_plane(double y, double limit, double ox, double oy, double oz, double dx, double dy, double dz,
                                            double& dist, double& px, double& py, double& pz) {
    if ((oy > y && dy > -K_1EM9) || (oy < y && dy < K_1EM9)) return false;
    double mult = (oy - y) * rcp64(-dy);
    if (mult > dist) return false;
    double x = fma(dx, mult, ox), yy = fma(dy, mult, oy), z = fma(dz, mult, oz);
    if (fabs(x) > limit || fabs(z) > limit) return false;  // NaN limit (unbounded) compares false
    dist = mult;
    px = x; py = yy; pz = z;
    return true;
}

// Unbounded plane against a ray whose direction d is NOT normalised (l2 = |d|^2); `dist` and the returned parameter are in
// units of |d|.  geometry.d:35-36 rejects dir.y > -1e-9 (origin above) / dir.y < 1e-9 (below) on the unit direction:
// here d.y has the wrong sign, or d.y^2 < 1e-18 |d|^2.
__device__ __forceinline__ bool isect_plane_u(double y, double ox, double oy, double oz, double dx, double dy, double dz, double l2,
                                              double& dist) {
    const bool grazing = dy * dy < 1e-18 * l2;
    if ((oy > y && (dy >= 0 || grazing)) || (oy < y && (dy <= 0 || grazing))) return false;
    double mult = (oy - y) * rcp64(-dy);
    if (mult > dist) return false;
    dist = mult;
    return true;
}

// geometry.d:92-125.  Every ray that reaches a geometry is unit (camera.d:144,172; scene.d:66-71; node.d:31-34 re-normalises in
// object space), so the reference's A = dot(d, d) is 1 to 2 ulp: the quadratic is solved in the half-b form
// t = -b -+ sqrt(b^2 - c), b = H . d, c = H . H - R^2 — the same roots as (-B -+ sqrt(B^2 - 4 A C)) / 2A to ~4e-16 relative,
// without the division and with a third fewer FP64 instructions.
__device__ __forceinline__ bool isect_sphere(const double* p, double ox, double oy, double oz, double dx, double dy, double dz,
                                             double& dist, double& px, double& py, double& pz) {
    double hx = ox - p[0], hy = oy - p[1], hz = oz - p[2];
    double b = dot3(hx, hy, hz, dx, dy, dz);
    double c = fma(-p[3], p[3], dot3(hx, hy, hz, hx, hy, hz));
    double D = fma(b, b, -c);
    if (D < 0) return false;
    double sq = D > 0 ? D * rsqrt64(D) : 0.0;
    double sol = -b - sq;
    if (sol < 0) sol = sq - b;
    if (sol < 0) return false;
    if (sol > dist) return false;
    dist = sol;
    px = fma(dx, sol, ox); py = fma(dy, sol, oy); pz = fma(dz, sol, oz);
    return true;
}

// one axis pass of geometry.d:199-235; (a) is the slab axis, (b, c) the in-face axes
__device__ __forceinline__ bool cube_pass(double oa, double ob, double oc, double da, double db, double dc, double ca, double cb,
                                          double cc, double half, double& dist, double& pa, double& pb, double& pc, int& side_out) {
    if (fabs(da) < K_1EM9) return false;
    bool found = false;
    const double inv = rcp64(-da);
#pragma unroll
    for (int side = -1; side <= 1; side += 2) {
        double mult = (oa - (ca + side * half)) * inv;
        if (mult < 0) continue;
        if (mult > dist) continue;
        double qb = fma(db, mult, ob), qc = fma(dc, mult, oc);
        if (qb < cb - half || qb > cb + half || qc < cc - half || qc > cc + half) continue;
        pa = fma(da, mult, oa); pb = qb; pc = qc;
        dist = mult;
        side_out = side > 0;
        found = true;
    }
    return found;
}

__device__ __forceinline__ bool isect_cube(const double* p, double ox, double oy, double oz, double dx, double dy, double dz,
                                           double& dist, double& px, double& py, double& pz, int& face) {
    double half = p[3] * 0.5;
    bool found = false;
    int side;
    if (cube_pass(oy, ox, oz, dy, dx, dz, p[1], p[0], p[2], half, dist, py, px, pz, side)) { found = true; face = 0 + side; }
    if (cube_pass(ox, oy, oz, dx, dy, dz, p[0], p[1], p[2], half, dist, px, py, pz, side)) { found = true; face = 2 + side; }
    if (cube_pass(oz, ox, oy, dz, dx, dy, p[2], p[0], p[1], half, dist, pz, px, py, side)) { found = true; face = 4 + side; }
    return found;
}

__device__ __forceinline__ bool isect_prim(const DevGeom& g, double ox, double oy, double oz, double dx, double dy, double dz,
                                           double& dist, double& px, double& py, double& pz, int& face) {
    if (g.type == C2RT_GEOM_PLANE) return isect_plane(g.p[0], g.p[1], ox, oy, oz, dx, dy, dz, dist, px, py, pz);
    if (g.type == C2RT_GEOM_SPHERE) return isect_sphere(g.p, ox, oy, oz, dx, dy, dz, dist, px, py, pz);
    return isect_cube(g.p, ox, oy, oz, dx, dy, dz, dist, px, py, pz, face);
}

__device__ __forceinline__ bool prim_inside(const DevGeom& g, double x, double y, double z) {
    if (g.type == C2RT_GEOM_SPHERE) {
        double ax = g.p[0] - x, ay = g.p[1] - y, az = g.p[2] - z;
        return dot3(ax, ay, az, ax, ay, az) < g.p[3] * g.p[3];
    }
    if (g.type == C2RT_GEOM_CUBE) {
        double h = g.p[3] * 0.5;
        return fabs(x - g.p[0]) <= h && fabs(y - g.p[1]) <= h && fabs(z - g.p[2]) <= h;
    }
    return false;  // Plane.isInside (geometry.d:25-28)
}

__device__ __forceinline__ bool csg_bool(int type, bool l, bool r) {  // geometry.d:361-364,371-374,399-402
    return type == C2RT_GEOM_CSG_UNION ? (l || r) : type == C2RT_GEOM_CSG_INTER ? (l && r) : (l && !r);
}

// CsgOp.isInside (geometry.d:334-337).  Depth-1 CSGs (primitive children) need no recursion; nested ones
// are evaluated by a depth-limited template recursion (NESTED_MAX_DEPTH levels, checked at scene create).
template <int D, bool BIG>
__device__ __forceinline__ bool geom_inside_d(int gi, double x, double y, double z) {
    const DevGeom& g = geom_at<BIG>(gi);
    if constexpr (D == 0) return prim_inside(g, x, y, z);
    else {
        if (g.type <= C2RT_GEOM_CUBE) return prim_inside(g, x, y, z);
        return csg_bool(g.type, geom_inside_d<D - 1, BIG>(g.left, x, y, z), geom_inside_d<D - 1, BIG>(g.right, x, y, z));
    }
}
template <bool BIG>
__device__ __forceinline__ bool geom_inside(int gi, double x, double y, double z) {
    const DevGeom& g = geom_at<BIG>(gi);
    if (g.type <= C2RT_GEOM_CUBE) return prim_inside(g, x, y, z);
    if (g.pad <= 1) return csg_bool(g.type, prim_inside(geom_at<BIG>(g.left), x, y, z), prim_inside(geom_at<BIG>(g.right), x, y, z));
    return geom_inside_d<3, BIG>(gi, x, y, z);
}

// ---------------------------------------------------------------- CSG
// CsgOp.intersect (geometry.d:292-332) for children that are convex primitives (nesting is rejected at
// scene-create time).  The reference collects each child's crossings by re-intersecting from
// p + d*1e-6 (findAllIntersections, geometry.d:271-290), shell-sorts the merged list
// (util/array.d:95-111) and walks it flipping inside-left / inside-right.  For a convex child that
// loop can only produce: nothing; the exit (origin inside); or entry + exit (origin outside) — and
// the distance it records for a child's 2nd crossing is 1e-6 SHORT of the true one, because the
// restart offset is never added back (geometry.d:283-286).  Both facts are reproduced here from the
// closed-form entry/exit distances, so the whole walk runs in registers: no restarts, no local arrays.
struct Crossings {
    double d0, d1;   // distances as the reference records them (d1 = true exit - 1e-6)
    int n, f0, f1;   // count (0..2) and cube-face codes
};

__device__ __forceinline__ Crossings cross_sphere(const double* p, double ox, double oy, double oz, double dx, double dy, double dz) {
    Crossings c;
    c.n = 0; c.f0 = 0; c.f1 = 0; c.d0 = 0; c.d1 = 0;
    double hx = ox - p[0], hy = oy - p[1], hz = oz - p[2];
    double b = dot3(hx, hy, hz, dx, dy, dz);          // half-b form for a unit direction (see isect_sphere)
    double cc = fma(-p[3], p[3], dot3(hx, hy, hz, hx, hy, hz));
    double D = fma(b, b, -cc);
    if (D < 0) return c;
    double sq = D > 0 ? D * rsqrt64(D) : 0.0;
    double x2 = -b - sq, x1 = sq - b;
    if (x2 >= 0) {
        c.d0 = x2;
        c.n = 1;
        // restart 1e-6 past the entry: the far root is found iff it is still ahead
        double rest = x1 - x2 - K_1EM6;
        if (rest >= 0) { c.d1 = x2 + rest; c.n = 2; }
    } else if (x1 >= 0) {
        c.d0 = x1;
        c.n = 1;
    }
    return c;
}

// entry / exit of the axis-aligned cube in the reference's face order (Y, X, Z passes; the later pass wins ties)
__device__ __forceinline__ Crossings cross_cube(const double* p, double ox, double oy, double oz, double dx, double dy, double dz) {
    Crossings c;
    c.n = 0; c.f0 = 0; c.f1 = 0; c.d0 = 0; c.d1 = 0;
    const double half = p[3] * 0.5;
    double tin = -K_1E300, tout = K_1E300;
    int fin = 0, fout = 0;
    bool miss = false;
    // pass order of geometry.d:172-191: Y (code 0), X (code 2), Z (code 4)
    const double o3[3] = {oy, ox, oz}, d3[3] = {dy, dx, dz}, c3[3] = {p[1], p[0], p[2]};
#pragma unroll
    for (int a = 0; a < 3; a++) {
        if (fabs(d3[a]) < K_1EM9) {
            // geometry.d:201-202 skips the pass; the other passes' in-face bounds still reject an origin outside this slab
            if (o3[a] < c3[a] - half || o3[a] > c3[a] + half) miss = true;
        } else {
            const double inv = rcp64(-d3[a]);
            const double tneg = (o3[a] - (c3[a] - half)) * inv;   // side -1 face
            const double tpos = (o3[a] - (c3[a] + half)) * inv;   // side +1 face
            const bool negfirst = tneg < tpos;
            const double tn = negfirst ? tneg : tpos, tf = negfirst ? tpos : tneg;
            if (tn >= tin) { tin = tn; fin = 2 * a + (negfirst ? 0 : 1); }
            if (tf <= tout) { tout = tf; fout = 2 * a + (negfirst ? 1 : 0); }
        }
    }
    if (miss || tin > tout || tout < 0) return c;
    if (tin >= 0) {
        c.d0 = tin; c.f0 = fin; c.n = 1;
        double rest = tout - tin - K_1EM6;
        if (rest >= 0) { c.d1 = tin + rest; c.f1 = fout; c.n = 2; }
    } else {
        c.d0 = tout; c.f0 = fout; c.n = 1;
    }
    return c;
}

__device__ __forceinline__ Crossings cross_plane(const double* p, double ox, double oy, double oz, double dx, double dy, double dz) {
    Crossings c;
    c.n = 0; c.f0 = 0; c.f1 = 0; c.d0 = 0; c.d1 = 0;
    double dist = K_1E99, px, py, pz;
    if (isect_plane(p[0], p[1], ox, oy, oz, dx, dy, dz, dist, px, py, pz)) { c.d0 = dist; c.n = 1; }
    return c;
}

__device__ __forceinline__ Crossings cross_prim(const DevGeom& g, double ox, double oy, double oz, double dx, double dy, double dz) {
    if (g.type == C2RT_GEOM_SPHERE) return cross_sphere(g.p, ox, oy, oz, dx, dy, dz);
    if (g.type == C2RT_GEOM_CUBE) return cross_cube(g.p, ox, oy, oz, dx, dy, dz);
    return cross_plane(g.p, ox, oy, oz, dx, dy, dz);
}

// compare-exchange with the strict `>` of IntersectionData.opCmp (intersectable.d:27-32)
__device__ __forceinline__ void cex(double& ka, int& ia, double& kb, int& ib) {
    if (ka > kb) {
        double t = ka; ka = kb; kb = t;
        int u = ia; ia = ib; ib = u;
    }
}

template <bool BIG>
__device__ __forceinline__ bool isect_csg(int gi, double ox, double oy, double oz, double dx, double dy, double dz, double& dist,
                                          double& px, double& py, double& pz, int& face, int& leaf) {
    const DevGeom& g = geom_at<BIG>(gi);
    const Crossings L = cross_prim(geom_at<BIG>(g.left), ox, oy, oz, dx, dy, dz);
    // inter / diff need the ray inside the LEFT child at the reported crossing (geometry.d:371-374,399-402); with no left
    // crossing inL stays false along the whole walk (L.n & 1 == 0), so nothing is reported: skip the right child
    if (g.type != C2RT_GEOM_CSG_UNION && L.n == 0) return false;
    const Crossings R = cross_prim(geom_at<BIG>(g.right), ox, oy, oz, dx, dy, dz);
    const int n = L.n + R.n;
    if (n == 0) return false;
    // merged list, left child's crossings first (geometry.d:301-302).  id: bit 1 = right child, bit 0 = second crossing,
    // bits 2.. = cube face code
    const double INF = CUDART_INF;
    double k0 = INF, k1 = INF, k2 = INF, k3 = INF;
    int i0 = 0, i1 = 0, i2 = 0, i3 = 0;
    {
        const int l0 = 0 | (L.f0 << 2), l1 = 1 | (L.f1 << 2), r0 = 2 | (R.f0 << 2), r1 = 3 | (R.f1 << 2);
        if (L.n == 2) {
            k0 = L.d0; i0 = l0; k1 = L.d1; i1 = l1;
            if (R.n >= 1) { k2 = R.d0; i2 = r0; }
            if (R.n == 2) { k3 = R.d1; i3 = r1; }
        } else if (L.n == 1) {
            k0 = L.d0; i0 = l0;
            if (R.n >= 1) { k1 = R.d0; i1 = r0; }
            if (R.n == 2) { k2 = R.d1; i2 = r1; }
        } else {
            k0 = R.d0; i0 = r0;
            if (R.n == 2) { k1 = R.d1; i1 = r1; }
        }
    }
    // util/array.d:95-111 for n <= 4: gap 2 exists only for n == 4 (two compare-exchanges), then the gap-1
    // insertion pass; padding slots hold +inf and never move
    if (n == 4) { cex(k0, i0, k2, i2); cex(k1, i1, k3, i3); }
    cex(k0, i0, k1, i1);
    cex(k1, i1, k2, i2); cex(k0, i0, k1, i1);
    cex(k2, i2, k3, i3); cex(k1, i1, k2, i2); cex(k0, i0, k1, i1);
    bool inL = L.n & 1, inR = R.n & 1;
    double kd = 0;
    int id = -1;
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const double ks = s == 0 ? k0 : s == 1 ? k1 : s == 2 ? k2 : k3;
        const int is = s == 0 ? i0 : s == 1 ? i1 : s == 2 ? i2 : i3;
        if (id < 0 && s < n) {
            if (is & 2) inR = !inR;
            else inL = !inL;
            if (csg_bool(g.type, inL, inR)) { kd = ks; id = is; }
        }
    }
    if (id < 0) return false;
    if (kd > dist) return false;
    dist = kd;
    const double tt = (id & 1) ? kd + K_1EM6 : kd;   // true parameter of the crossing point
    px = fma(dx, tt, ox); py = fma(dy, tt, oy); pz = fma(dz, tt, oz);
    face = id >> 2;
    leaf = (id & 2) ? g.right : g.left;
    if (g.type == C2RT_GEOM_CSG_DIFF) {
        const DevGeom& gr = geom_at<BIG>(g.right);   // a primitive: this closed form is only used for CSGs of primitives
        bool a = prim_inside(gr, px - dx * K_1EM6, py - dy * K_1EM6, pz - dz * K_1EM6);
        bool b = prim_inside(gr, px + dx * K_1EM6, py + dy * K_1EM6, pz + dz * K_1EM6);
        if (a != b) face |= FACE_FLIP;
    }
    return true;
}

// ---------------------------------------------------------------- nested CSG (literal emulation)
// A CSG whose child is itself a CSG has no closed form: the child can yield many crossings, and the
// reference's walk has observable quirks there — a nested child's crossings carry the LEAF primitive in
// `g`, so `current.g is left` is false and they toggle the RIGHT flag (geometry.d:314), and
// a CSG reports interior boundaries as hits while the boolean stays true (geometry.d:321).  This path
// therefore replays findAllIntersections / sort / walk literally, recursing through a depth-limited
// template; it is only compiled into the MODE_NESTED kernel.
constexpr int NESTED_MAX_CROSSINGS = 8;   // per child
constexpr int NESTED_MAX_DEPTH = 3;       // CSG levels above the primitives (validated at scene create)

struct LitCrossing {
    double dist, px, py, pz;
    int face, leaf;
};

template <int D, bool BIG>
__device__ __noinline__ bool isect_geom_lit(int gi, double ox, double oy, double oz, double dx, double dy, double dz, double& dist,
                                            double& px, double& py, double& pz, int& face, int& leaf);

template <int D, bool BIG>
__device__ __forceinline__ int find_all_lit(int gi, double ox, double oy, double oz, double dx, double dy, double dz, LitCrossing* out) {
    double cur = 0;
    int n = 0;
    while (n < NESTED_MAX_CROSSINGS) {  // geometry.d:271-290
        double dist = 1e99, px, py, pz;
        int face = 0, leaf = gi;
        if (!isect_geom_lit<D - 1, BIG>(gi, ox, oy, oz, dx, dy, dz, dist, px, py, pz, face, leaf)) break;
        dist += cur;
        cur = dist;
        ox = fma(dx, 1e-6, px); oy = fma(dy, 1e-6, py); oz = fma(dz, 1e-6, pz);
        out[n].dist = dist; out[n].px = px; out[n].py = py; out[n].pz = pz;
        out[n].face = face; out[n].leaf = leaf;
        n++;
    }
    return n;
}

template <int D, bool BIG>
__device__ __noinline__ bool isect_geom_lit(int gi, double ox, double oy, double oz, double dx, double dy, double dz, double& dist,
                                            double& px, double& py, double& pz, int& face, int& leaf) {
    const DevGeom& g = geom_at<BIG>(gi);
    if (g.type <= C2RT_GEOM_CUBE) {
        leaf = gi;
        face = 0;
        return isect_prim(g, ox, oy, oz, dx, dy, dz, dist, px, py, pz, face);
    }
    if constexpr (D == 0) return false;  // deeper than NESTED_MAX_DEPTH: rejected at scene create
    else {
    LitCrossing all[2 * NESTED_MAX_CROSSINGS];
    const int nl = find_all_lit<D, BIG>(g.left, ox, oy, oz, dx, dy, dz, all);
    const int nr = find_all_lit<D, BIG>(g.right, ox, oy, oz, dx, dy, dz, all + nl);
    const int n = nl + nr;
    // util/array.d:95-111 shell sort, including the `ref` loop index and the gap sequence
    int inc = n / 2;
    while (inc) {
        for (int key = 0; key < n; key++) {
            int i = key;
            LitCrossing elem = all[i];
            while (i >= inc && all[i - inc].dist > elem.dist) {
                all[i] = all[i - inc];
                i -= inc;
            }
            all[i] = elem;
            key = i;
        }
        inc = (inc == 2) ? 1 : (int)(inc * 5.0 / 11);
    }
    bool inL = nl & 1, inR = nr & 1;
    for (int k = 0; k < n; k++) {
        if (all[k].leaf == g.left) inL = !inL;   // `current.g is left`: only true when the left child is a primitive
        else inR = !inR;
        if (csg_bool(g.type, inL, inR)) {
            if (all[k].dist > dist) return false;
            dist = all[k].dist;
            px = all[k].px; py = all[k].py; pz = all[k].pz;
            face = all[k].face;
            leaf = all[k].leaf;
            if (g.type == C2RT_GEOM_CSG_DIFF) {
                bool a = geom_inside<BIG>(g.right, px - dx * 1e-6, py - dy * 1e-6, pz - dz * 1e-6);
                bool b = geom_inside<BIG>(g.right, px + dx * 1e-6, py + dy * 1e-6, pz + dz * 1e-6);
                if (a != b) face ^= FACE_FLIP;
            }
            return true;
        }
    }
    return false;
    }
}

// ---------------------------------------------------------------- node
// Conservative FP32 bounding-sphere rejection.  Result-identical: it only skips nodes the exact
// FP64 test would reject.  `margin`/`slack` bound the FP32 rounding of everything in the test
// (|err(d2)| <= 24 eps S^2, |err(tca)| <= 5 eps S with S = |centre| + |origin|).
__device__ __forceinline__ bool cull_sphere(float bx, float by, float bz, float br, float br2, float bclen, const Ray& r, float tmaxf) {
    float cx = bx - r.fox, cy = by - r.foy, cz = bz - r.foz;
    float tca = dot3f(cx, cy, cz, r.fdx, r.fdy, r.fdz);
    float c2 = dot3f(cx, cy, cz, cx, cy, cz);
    float S = bclen + r.olen;
    float margin = 4e-6f * S * S;
    float slack = 1e-6f * S;
    float d2 = fmaf(-tca, tca, c2);
    if (d2 > br2 + margin) return true;         // the line misses the sphere
    if (tca + br < -slack) return true;         // the sphere is behind the origin
    if (tca - br > tmaxf + slack) return true;  // the sphere starts beyond the best distance so far
    return false;
}
__device__ __forceinline__ bool cull(const DevNode& nd, const Ray& r, float tmaxf) {
    return cull_sphere(nd.bcf[0], nd.bcf[1], nd.bcf[2], nd.brf, nd.br2f, nd.bclen, r, tmaxf);
}

// Exact FP64 test of node `ni` for the scene classes with bounded / generic nodes: node.d:23-49 + transform.d:57-86 around the
// primitive / CSG test.  ONE copy serves camera and shadow rays (trace_warp), and one copy of each primitive test serves the
// world-space fast-path kinds (identity-transform primitives and diagonally scaled unbounded planes, tested against their
// pre-offset parameters nd.wp with the world ray) and the object-space path: the kernel's instruction footprint is what
// its issue rate hangs on.  Every branch on nd / g fields is warp-uniform (the node index is).
// Returns true and updates `h` iff the node yields a hit with dist <= h.dist; h.p is in the node's local frame.
template <int MODE>
__device__ __forceinline__ bool node_hit(int ni, const DevNode& nd, const Ray& r, HitRec& h) {
    constexpr bool BIG = is_big(MODE);
    const DevGeom& g = geom_at<BIG>(nd.geom);
    // (a scene class without MODE_GENERIC holds world-space fast-path nodes only: the object-space code compiles away)
    const bool world = !(MODE & MODE_GENERIC) || nd.kind != KIND_GENERIC;
    const bool plain = world || (nd.flags & NODE_IDENTITY);   // no matrix: at most an offset
    double ox = r.ox, oy = r.oy, oz = r.oz, dx = r.dx, dy = r.dy, dz = r.dz, len = 1.0;
    if (!world) { ox -= nd.off[0]; oy -= nd.off[1]; oz -= nd.off[2]; }
    if (!plain) {
        const double tx = ox, ty = oy, tz = oz;
        mulvm(nd.Minv, tx, ty, tz, ox, oy, oz);
        mulvm(nd.Minv, r.dx, r.dy, r.dz, dx, dy, dz);
        const double l2 = dot3(dx, dy, dz, dx, dy, dz);
        const double inv = rsqrt64(l2);
        len = l2 * inv;
        dx *= inv; dy *= inv; dz *= inv;
    }
    double dist = plain ? h.dist : h.dist * len;
    double px, py, pz;
    int face = 0, leaf = nd.geom;
    const double* prm = world ? nd.wp : g.p;
    bool hit;
    if (g.type == C2RT_GEOM_PLANE) hit = isect_plane(prm[0], world ? CUDART_NAN : prm[1], ox, oy, oz, dx, dy, dz, dist, px, py, pz);
    else if (g.type == C2RT_GEOM_SPHERE) hit = isect_sphere(prm, ox, oy, oz, dx, dy, dz, dist, px, py, pz);
    else if (g.type == C2RT_GEOM_CUBE) hit = isect_cube(prm, ox, oy, oz, dx, dy, dz, dist, px, py, pz, face);
    else if (!(MODE & MODE_GENERIC)) hit = false;   // (CSG nodes are KIND_GENERIC)
    else if ((MODE & MODE_NESTED) && g.pad != 1) hit = isect_geom_lit<NESTED_MAX_DEPTH, BIG>(nd.geom, ox, oy, oz, dx, dy, dz, dist, px, py, pz, face, leaf);
    else hit = isect_csg<BIG>(nd.geom, ox, oy, oz, dx, dy, dz, dist, px, py, pz, face, leaf);
    if (!hit) return false;
    h.dist = plain ? dist : dist * rcp64(len);
    h.px = px; h.py = py; h.pz = pz;
    h.node = ni; h.leaf = leaf; h.face = face;
    return true;
}

// MODE_SOLO frames without DOF / stereo whose geometry is REGULAR (fp.solo_fast, decided per frame by c2rt_api.cu fill_params;
// every other frame of the scene class runs on the general sampling kernel): every camera ray starts at the camera position,
// off the scene's one plane, so
//  * which side of the plane it starts on and its height above it are frame constants (fp.solo_sign, fp.solo_h), and
//    geometry.d:35-36 reduces to the SIGN BIT of d.y (an integer test on its high word) plus the grazing test
//    d.y^2 < 1e-18 |d|^2 — which cannot hold while d.y^2 >= fp.graze_dy2 = 1e-18 max|d|^2 over the frame's rays, so |d|^2 is
//    only formed for the (practically absent) rays below that: out of line, or the compiler hoists the dot product above
//    the test.  (d.y = -0.0 / +0.0 pass the sign test on the "wrong" side and are rejected as grazing, like the reference does.)
//  * the hit distance |h| / |d.y| <= |h| 1e9 / min|d| stays far below the initial 1e99 (geometry.d:39 never rejects);
//  * every hit faces the camera's side (faceforward, imported_types.d:69-73), the shadow-ray origin p + N 1e-6 lies strictly
//    on that side of the plane, and so does the light, by more than 1e-5: geometry.d:35-36 rejects the shadow ray by side
//    and sign, testVisibility (scene.d:62-78) finds nothing — the one plane cannot shadow itself.
__device__ __noinline__ bool grazing_exact(double dx, double dy, double dz) { return dy * dy < 1e-18 * dot3(dx, dy, dz, dx, dy, dz); }
// `h`: the ray origin's height above the plane (fp.solo_h for a fixed camera; per ray under DOF / stereo, where fill_params has
// checked the whole lens against the same conditions)
__device__ __forceinline__ bool isect_plane_solo(const FrameParams& fp, double h, double dx, double dy, double dz, double& dist) {
    if ((int)((unsigned)__double2hiint(dy) ^ fp.solo_sign) < 0) return false;   // d.y points away from the plane
    if (dy * dy < fp.graze_dy2) {
        if (grazing_exact(dx, dy, dz)) return false;
    }
    dist = h * rcp64(-dy);
    return true;
}
// the kernels without the sampling loop only run regular frames; the sampling kernels run both kinds (fp.solo_fast, warp-uniform)
__host__ __device__ constexpr bool solo_fast(int mode) { return (mode & MODE_SOLO) && !(mode & MODE_SAMPLING); }
template <int MODE>
__device__ __forceinline__ bool solo_regular(const FrameParams& fp) {
    if constexpr (solo_fast(MODE)) return true;
    else if constexpr ((MODE & MODE_SOLO) != 0) return fp.solo_fast != 0;
    else return false;
}

// Plane-only scene classes: no bounded and no generic node exists, i.e. every node is a world-space plane.
// CAMERA_RAY: `r` comes from gen_ray (un-normalised there); shadow rays are always unit.  The hit point is o + d * dist,
// computed for the winning hit only (surface_of).
template <int MODE, bool CAMERA_RAY>
__device__ __forceinline__ bool node_exact(int ni, const DevNode& nd, const Ray& r, HitRec& h) {
    static_assert(plane_only(MODE), "scene classes with bounded / generic nodes go through node_hit");
    bool hit;
    double px, py, pz;
    if (CAMERA_RAY) hit = isect_plane_u(nd.wp[0], r.ox, r.oy, r.oz, r.dx, r.dy, r.dz, r.l2, h.dist);
    else hit = isect_plane(nd.wp[0], CUDART_NAN, r.ox, r.oy, r.oz, r.dx, r.dy, r.dz, h.dist, px, py, pz);
    if (hit) { h.node = ni; h.leaf = nd.geom; h.face = 0; }
    return hit;
}

// ---------------------------------------------------------------- per-warp node masks (warp ballot / vote)
// The reference tests every ray against every node (renderer.d:336-338, scene.d:73-75).  Here the 32 lanes of a warp
// (an 8x4 pixel patch) first decide TOGETHER which nodes any of their rays can reach: lane l tests the bounding spheres of
// nodes l, l + 32, ... against a volume that contains all 32 rays, and one __ballot_sync per 32 nodes turns the answers
// into a bit mask that every lane then walks in scene order (so the reference's closest-hit and tie rules,
// geometry.d:43,111,214, are untouched).  The walk is warp-uniform — same node for all lanes, per-lane work predicated —
// and a shadow walk is left through __all_sync as soon as every lane that needs an answer has found an occluder
// (the reference's early return at the first occluder, scene.d:73-75, taken by the whole warp at once).
static_assert(C2RT_MAX_NODES <= 64, "a node mask of the constant-memory scene block is two 32-bit words");
constexpr unsigned FULL_WARP = 0xffffffffu;
constexpr int WARPS_PER_CTA = BLOCK_THREADS / 32;
constexpr int BIG_MASK_WORDS = C2RT_MAX_NODES_GLOBAL / 32;
typedef uint32_t NodeWord;
// A warp's node mask: one bit per node, 32 nodes per word (one ballot each).  Constant-block scenes (<= 64 nodes): two words in
// (uniform) registers.  MODE_BIG scenes: the words live in shared memory, a row per warp (`more`).
struct NodeMask {
    NodeWord w0, w1;
    NodeWord* more;
};
__device__ __forceinline__ NodeWord all_nodes_word(int first) {
    const int left = c_scene.n_nodes - first;
    return left >= 32 ? 0xffffffffu : left > 0 ? (1u << left) - 1u : 0u;
}
__device__ __forceinline__ int mask_words() { return (c_scene.n_nodes + 31) >> 5; }
template <bool BIG>
__device__ __forceinline__ NodeWord mask_word(const NodeMask& m, int k) { return BIG ? m.more[k] : (k ? m.w1 : m.w0); }
// lane l answers for node 32 k + l; one ballot makes word k
template <bool BIG, class F>
__device__ __forceinline__ void ballot_nodes(NodeMask& m, F&& reaches) {
    const int lane = (int)(threadIdx.x & 31u), n = c_scene.n_nodes;
    m.w0 = 0; m.w1 = 0;
#pragma unroll 1
    for (int k = 0; k < mask_words(); k++) {
        const int i = 32 * k + lane;
        const NodeWord w = __ballot_sync(FULL_WARP, i < n && reaches(i));
        if (BIG) { if (lane == 0) m.more[k] = w; }
        else if (k == 0) m.w0 = w;
        else m.w1 = w;
    }
    if (BIG) __syncwarp();
}
template <bool BIG>
__device__ __forceinline__ void all_nodes_mask(NodeMask& m) {
    m.w0 = all_nodes_word(0);
    m.w1 = all_nodes_word(32);
    if (BIG) {
        const int lane = (int)(threadIdx.x & 31u);
        for (int k = lane; k < mask_words(); k += 32) m.more[k] = all_nodes_word(32 * k);
        __syncwarp();
    }
}
// warp-wide FP32 min / max in one instruction each (CREDUX, sm_100a), result in a uniform register
__device__ __forceinline__ float warp_min(float v) {
    float r;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float warp_max(float v) {
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}

// Which nodes can a shadow ray of this warp reach?  Every shadow segment runs from a lane's origin to the light, so all of
// them lie in the convex hull of {light} and the lanes' origins, which lies in the capsule around the segment
// (centre of the origins' bounding box -> light) with the box's half-diagonal as radius.  A node whose bounding sphere
// stays clear of that capsule cannot occlude any lane.  FP32 with an explicit rounding margin (8e-6 of the magnitudes
// involved: the arithmetic below loses < 1e-6 of them); the node spheres are already inflated (c2rt_api.cu).
template <int MODE>
__device__ __forceinline__ void shadow_mask(NodeMask& m, const float4* __restrict__ bounds, bool need, const Ray& r, const DevLight& L, bool& any) {
    constexpr bool BIG = is_big(MODE);
    bool masked = (MODE & MODE_BOUNDED) != 0;
#ifdef C2RT_NO_WARP_MASK
    masked = false;
#endif
    if (!masked) {
        any = __any_sync(FULL_WARP, need);
        all_nodes_mask<BIG>(m);
        return;
    }
    const float INF = __int_as_float(0x7f800000);
    const float mnx = warp_min(need ? r.fox : INF), mxx = warp_max(need ? r.fox : -INF);
    const float mny = warp_min(need ? r.foy : INF), mxy = warp_max(need ? r.foy : -INF);
    const float mnz = warp_min(need ? r.foz : INF), mxz = warp_max(need ? r.foz : -INF);
    any = mnx <= mxx;
    if (!any) return;   // no lane of this warp needs a shadow ray (warp-uniform); the caller skips the walk
    const float cx = 0.5f * (mnx + mxx), cy = 0.5f * (mny + mxy), cz = 0.5f * (mnz + mxz);
    const float hx = mxx - cx, hy = mxy - cy, hz = mxz - cz;
    const float rho = sqrt_up(dot3f(hx, hy, hz, hx, hy, hz));
    const float dx = L.posf[0] - cx, dy = L.posf[1] - cy, dz = L.posf[2] - cz;   // capsule axis: box centre -> light
    const float dd = dot3f(dx, dy, dz, dx, dy, dz);
    const float inv_dd = dd > 0.f ? __fdividef(1.0f, dd) : 0.f;   // (approximate: t only picks the point of the axis; the margin covers it)
    const float mag = fabsf(cx) + fabsf(cy) + fabsf(cz) + fabsf(dx) + fabsf(dy) + fabsf(dz);   // L1 >= L2: a cheaper, larger margin
    ballot_nodes<BIG>(m, [&](int i) {
        // per-lane index: the bound comes from GLOBAL memory (one coalesced 16-byte load per lane) — a constant-bank read
        // with 32 different addresses would be replayed 32 times
        const float4 b = __ldg(&bounds[i]);
        if (b.w < 0.f) return true;   // unbounded (planes)
        const float ax = b.x - cx, ay = b.y - cy, az = b.z - cz;
        const float t = fminf(fmaxf(dot3f(ax, ay, az, dx, dy, dz) * inv_dd, 0.f), 1.f);
        const float qx = fmaf(-t, dx, ax), qy = fmaf(-t, dy, ay), qz = fmaf(-t, dz, az);
        const float reachr = b.w + rho + 8e-6f * (mag + fabsf(b.x) + fabsf(b.y) + fabsf(b.z) + rho);
        return !(dot3f(qx, qy, qz, qx, qy, qz) > reachr * reachr);   // (a NaN bound keeps the node)
    });
}

// testVisibility for the plane-only scene classes.  Dy = light.y - from.y in FP64 settles almost every plane by sign;
// the full FP64 shadow ray (scene.d:66-71) is only built for a plane that lies between the two heights.
template <int MODE>
__device__ __forceinline__ bool occluded_planes(double fx, double fy, double fz, const DevLight& L, double Dy) {
    bool exact = false;
    Ray r;
    HitRec h;
    const int n = (MODE & MODE_SOLO) ? 1 : c_scene.n_nodes;
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        const DevNode& nd = c_scene.nodes[(MODE & MODE_SOLO) ? 0 : i];
        const double y = nd.wp[0];
        if ((fy > y && Dy >= 0) || (fy < y && Dy <= 0)) continue;   // implied by geometry.d:35-36
        if (!exact) {
            const double Dx = L.pos[0] - fx, Dz = L.pos[2] - fz;
            const double len2 = dot3(Dx, Dy, Dz, Dx, Dy, Dz);
            const double inv = rsqrt64(len2);
            r.ox = fx; r.oy = fy; r.oz = fz;
            r.dx = Dx * inv; r.dy = Dy * inv; r.dz = Dz * inv;
            h.dist = len2 * inv;
            exact = true;
        }
        if (node_exact<MODE, false>(i, nd, r, h)) return true;
    }
    return false;
}

// ---------------------------------------------------------------- textures
__device__ __forceinline__ int cast_int_x86(double v) {  // cvttsd2si: out of range / NaN -> INT_MIN
    if (!(v > -2147483649.0 && v < 2147483648.0)) return (int)0x80000000;
    return (int)v;
}

// sin(2 pi u f) for FP64 u and f = frequency / 2 pi (folded at scene create): the product is reduced to a
// fraction of a revolution in FP64 — fma(u, f, MAGIC) rounds u f to the nearest integer k, a second fma
// gives u f - k with a single rounding — then the FP32 SFU sine (|abs err| < 2e-6).  The reference takes
// sin(u * frequency) in FP64 and narrows to float (texture.d:82-83).  The fraction is biased by +1 into
// [0.5, 1.5] so its FP64 -> FP32 conversion can be a truncating bit shuffle on the ALU pipe: F2F shares the
// quarter-rate XU pipe with MUFU.SIN, which this texture keeps busy.
__device__ __forceinline__ float sin_rev(double u, double f) {
    const double MAGIC = 6755399441055744.0;                      // 1.5 * 2^52
    const double k1 = fma(u, f, MAGIC) - (MAGIC + 1.0);           // nearest integer - 1, exact
    const double fr = fma(u, f, -k1);
    const unsigned hi = (unsigned)__double2hiint(fr), lo = (unsigned)__double2loint(fr);
    const float x = __uint_as_float(__funnelshift_l(lo, hi - 0x38000000u, 3));   // exponent rebias 1023 -> 127, top 23 mantissa bits
    return __sinf(x * 6.2831853f);
}

// The same sine with the range reduction in ONE FP64 instruction: F = f * 2^32 (DevTex.d holds that), so fma(u, F, 1.5 * 2^52)
// leaves round(u f 2^32) mod 2^32 — the phase as a 32-bit fraction of a revolution, wrap-around included — in the low mantissa
// word.  Its top 23 bits become the mantissa of a float in [1, 2) (whole revolutions do not matter to a sine): the same 2^-23
// revolution granularity as sin_rev, with one DFMA where sin_rev spends two DFMA and a DADD.  Valid while |u f| < 2^19
// revolutions (DevTex.w / .h hold the |u| / |v| that keep it under 2^18, as the high word of a double); sample_texture falls back
// to sin_rev beyond.
__device__ __forceinline__ float sin_phase(double u, double F) {
    const unsigned p = (unsigned)__double2loint(fma(u, F, 6755399441055744.0));
    return __sinf(__uint_as_float((p >> 9) | 0x3f800000u) * 6.2831853f);
}

// MODE_SOLO kernels compiled for a bitmap texture stage its 4 KB palette (if it has one) in shared memory once per CTA: the
// palette read depends on the index-quad load, and a shared-memory read is the shorter second hop (C3: the zaphod page)
__host__ __device__ constexpr bool solo_bitmap(int mode) {
    return (mode & MODE_SOLO) && ((mode & MODE_TEX_MASK) >> MODE_TEX_SHIFT) == 1 + C2RT_TEX_BITMAP;
}
__shared__ float4 s_solo_palette[256];

// Bitmap.getFilteredPixel (bitmap.d:48-63): bilinear fetch at float texel coordinates (x, y), wrapping at the edges;
// out-of-range coordinates (incl. NaN) give NamedColors.red
// `spal`: the texture's palette staged in shared memory (MODE_SOLO bitmap kernels, render_frame_kernel), or nullptr
__device__ __forceinline__ Col bitmap_fetch(const DevTex& t, float x, float y, const float4* spal = nullptr) {
    if (!(x >= 0.f) || !(y >= 0.f) || !(x < (float)t.w) || !(y < (float)t.h))   // (t.w, t.h are integers: x < w <=> (size_t)x < w)
        return mkcol(1.f, 0.f, 0.f);
    int tx = (int)x, ty = (int)y;
    int txn = tx + 1 == t.w ? 0 : tx + 1, tyn = ty + 1 == t.h ? 0 : ty + 1;
    float p = x - (float)tx, q = y - (float)ty;
    float4 a, b, c, d;
    if (t.quads) {   // palette form (scene_dev.h DevTex): one 4-byte load, then the 4 KB palette (shared memory or L1)
        const uint32_t k = __ldg(&t.quads[(size_t)ty * t.w + tx]);
        if (spal) {
            a = spal[k & 255u]; b = spal[(k >> 8) & 255u]; c = spal[(k >> 16) & 255u]; d = spal[k >> 24];
        } else {
            a = __ldg(&t.palette[k & 255u]);
            b = __ldg(&t.palette[(k >> 8) & 255u]);
            c = __ldg(&t.palette[(k >> 16) & 255u]);
            d = __ldg(&t.palette[k >> 24]);
        }
    } else {
        a = __ldg(&t.texels[(size_t)ty * t.w + tx]);
        b = __ldg(&t.texels[(size_t)ty * t.w + txn]);
        c = __ldg(&t.texels[(size_t)tyn * t.w + tx]);
        d = __ldg(&t.texels[(size_t)tyn * t.w + txn]);
    }
    float wa = (1.0f - p) * (1.0f - q), wb = p * (1.0f - q), wc = (1.0f - p) * q, wd = p * q;
    return mkcol(a.x * wa + b.x * wb + c.x * wc + d.x * wd, a.y * wa + b.y * wb + c.y * wc + d.y * wd,
                 a.z * wa + b.z * wb + c.z * wc + d.z * wd);
}

// Texture lookup at (u, v): texture.d:36-54 (Checker), :77-86 (Procedure2), :116-126 + bitmap.d:48-63 (bitmap, bilinear)
// MODE_SOLO: the texture is record 0 and its kind is a template constant
template <int MODE>
__device__ __forceinline__ Col sample_texture(int ti, double u, double v) {
    constexpr bool SOLO = (MODE & MODE_SOLO) != 0;
    const DevTex& t = tex_at<is_big(MODE)>(SOLO ? 0 : ti);
    const int type = SOLO ? ((MODE & MODE_TEX_MASK) >> MODE_TEX_SHIFT) - 1 : t.type;
    if (type == C2RT_TEX_CHECKER) {
        int x = cast_int_x86(floor(u * t.d[1]));
        int y = cast_int_x86(floor(v * t.d[1]));
        int white = (int)((unsigned)x + (unsigned)y) % 2;
        return white ? mkcol(t.c[3], t.c[4], t.c[5]) : mkcol(t.c[0], t.c[1], t.c[2]);
    }
    if (type == C2RT_TEX_PROCEDURE2) {
        // six sines (texture.d:82-83)
        float su[3], sv[3];
        const bool near_u = ((unsigned)__double2hiint(u) & 0x7fffffffu) < (unsigned)t.w;
        const bool near_v = ((unsigned)__double2hiint(v) & 0x7fffffffu) < (unsigned)t.h;
        if (near_u && near_v) {
#pragma unroll
            for (int i = 0; i < 3; i++) { su[i] = sin_phase(u, t.d[i]); sv[i] = sin_phase(v, t.d[3 + i]); }
        } else {   // beyond 2^18 revolutions (a far horizon, or NaN): the two-step reduction
            const double R = 2.3283064365386962890625e-10;   // 2^-32, exact
#pragma unroll
            for (int i = 0; i < 3; i++) { su[i] = sin_rev(u, t.d[i] * R); sv[i] = sin_rev(v, t.d[3 + i] * R); }
        }
        // (the sum starts from its first term, not from 0.f + term: the same value unless the term is -0)
        Col res = mkcol(t.c[0] * su[0] + t.c[9] * sv[0], t.c[1] * su[0] + t.c[10] * sv[0], t.c[2] * su[0] + t.c[11] * sv[0]);
#pragma unroll
        for (int i = 1; i < 3; i++) {
            res.r += t.c[3 * i + 0] * su[i] + t.c[9 + 3 * i + 0] * sv[i];
            res.g += t.c[3 * i + 1] * su[i] + t.c[9 + 3 * i + 1] * sv[i];
            res.b += t.c[3 * i + 2] * su[i] + t.c[9 + 3 * i + 2] * sv[i];
        }
        return res;
    }
    // bitmap: texture.d:116-126 + bitmap.d:48-63
    u *= t.d[0];
    v *= t.d[0];
    u = u - floor(u);
    v = v - floor(v);
    return bitmap_fetch(t, (float)u * (float)t.w, (float)v * (float)t.h, solo_bitmap(MODE) ? s_solo_palette : nullptr);
}

// environment.d:7-10 for a miss (renderer.d:366-368): black, or — EXTENSION, no counterpart in the reference (c2rt.h
// C2RT_ENV_CUBEMAP, oracle/orc_scene.hpp Environment) — the bilinear sample of the cube face the direction's largest component
// points at.  (dx, dy, dz) need not be unit: only ratios of its components are used.
__device__ __noinline__ Col env_lookup(double dx, double dy, double dz) {
    if (c_scene.env_type != C2RT_ENV_CUBEMAP) return mkcol(0.f, 0.f, 0.f);
    const double ax = fabs(dx), ay = fabs(dy), az = fabs(dz);
    int face;
    double sx, sy;   // face coordinates in [-1, 1]
    if (ax >= ay && ax >= az) {
        if (!(ax > 0)) return mkcol(0.f, 0.f, 0.f);
        const double inv = rcp64(ax), vy = dy * inv, vz = dz * inv;
        face = dx < 0 ? 1 : 0;
        sx = dx < 0 ? vz : -vz; sy = -vy;
    } else if (ay >= az) {
        const double inv = rcp64(ay), vx = dx * inv, vz = dz * inv;
        face = dy < 0 ? 3 : 2;
        sx = vx; sy = dy < 0 ? -vz : vz;
    } else {
        const double inv = rcp64(az), vx = dx * inv, vy = dy * inv;
        face = dz < 0 ? 5 : 4;
        sx = dz < 0 ? -vx : vx; sy = -vy;
    }
    const DevTex& t = c_scene.env_faces[face];
    return bitmap_fetch(t, (float)((sx + 1.0) * 0.5 * (double)(t.w - 1)), (float)((sy + 1.0) * 0.5 * (double)(t.h - 1)));
}

// ---------------------------------------------------------------- hit completion + shading
struct Surface {
    double px, py, pz;     // world-space hit point
    double gx, gy, gz;     // world-space geometric normal, NOT normalised: only its sign against the ray is taken in FP64
    float nx, ny, nz;      // unit world-space normal for lighting
    double u, v;           // texture coordinates
};

// Local-frame normal / uv from (leaf, face, p): geometry.d:49-55 (plane), :114-120 (sphere), :224-230 (cube).
// `c` are the primitive's parameters in the frame of h.p.
__device__ __forceinline__ void local_surface(int type, const double* c, const HitRec& h, bool need_uv, double& gx, double& gy, double& gz,
                                              double& u, double& v) {
    u = 0; v = 0;
    if (type == C2RT_GEOM_PLANE) {
        gx = 0; gy = 1; gz = 0;
        u = h.px; v = h.pz;
    } else if (type == C2RT_GEOM_SPHERE) {
        gx = h.px - c[0]; gy = h.py - c[1]; gz = h.pz - c[2];
        if (need_uv) {
            // sphere uv in FP32 from the FP64 difference vector; the seam is kept inside [-pi, pi]
            float fx = (float)gx, fy = (float)gy, fz = (float)gz;
            float angle = atan2f(fz, fx);
            double ad = fmin(fmax((double)angle, -CUDART_PI), CUDART_PI);
            u = (CUDART_PI + ad) * (0.5 / CUDART_PI);
            float as = atan2f(fy, sqrtf(fmaf(fx, fx, fz * fz)));  // asin(dy / R) without the pole singularity
            v = 1.0 - (CUDART_PI / 2 + (double)as) * (1.0 / CUDART_PI);
        }
    } else {
        int axis = (h.face & 7) >> 1;
        double s = (h.face & 1) ? 1.0 : -1.0;
        gx = axis == 1 ? s : 0.0;
        gy = axis == 0 ? s : 0.0;
        gz = axis == 2 ? s : 0.0;
        // u, v stay in the permuted frame of the pass that produced the hit (quirk)
        if (axis == 0) { u = h.px - c[0]; v = h.pz - c[2]; }
        else if (axis == 1) { u = h.py - c[1]; v = h.pz - c[2]; }
        else { u = h.px - c[0]; v = h.py - c[1]; }
    }
    if (h.face & FACE_FLIP) { gx = -gx; gy = -gy; gz = -gz; }
}

// `ray` == nullptr: hin.p is already complete (debug pixel pick)
template <int MODE>
__device__ __forceinline__ void surface_of(const HitRec& hin, const Ray* ray, bool need_uv, Surface& s) {
    const DevNode& nd = node_at<is_big(MODE)>((MODE & MODE_SOLO) ? 0 : hin.node);
    HitRec h = hin;
    if (plane_only(MODE)) {   // every node is a world-space plane: normal (0, 1, 0), uv = object-space x, z (geometry.d:49-55)
        if (ray) { h.px = fma(ray->dx, h.dist, ray->ox); h.py = fma(ray->dy, h.dist, ray->oy); h.pz = fma(ray->dz, h.dist, ray->oz); }
        s.gx = 0; s.gy = 1; s.gz = 0;
        s.u = (h.px - nd.off[0]) * nd.wp[1];
        s.v = (h.pz - nd.off[2]) * nd.wp[2];
        s.px = h.px; s.py = h.py; s.pz = h.pz;
        s.nx = 0.f; s.ny = 1.f; s.nz = 0.f;
        return;
    }
    const DevGeom& g = geom_at<is_big(MODE)>(hin.leaf);
    if (nd.kind != KIND_GENERIC) {
        // world-space fast path: the hit point is o + d * dist (exactly what the intersector computed), nd.wp the pre-offset parameters
        if (ray) { h.px = fma(ray->dx, h.dist, ray->ox); h.py = fma(ray->dy, h.dist, ray->oy); h.pz = fma(ray->dz, h.dist, ray->oz); }
        local_surface(g.type, nd.wp, h, need_uv, s.gx, s.gy, s.gz, s.u, s.v);
        if (nd.kind == KIND_PLANE_W) {  // uv are object-space (geometry.d:54-55): undo the offset and the diagonal scale
            s.u = (h.px - nd.off[0]) * nd.wp[1];
            s.v = (h.pz - nd.off[2]) * nd.wp[2];
        }
        s.px = h.px; s.py = h.py; s.pz = h.pz;
    } else {
        double gx, gy, gz;
        local_surface(g.type, g.p, h, need_uv, gx, gy, gz, s.u, s.v);
        if (nd.flags & NODE_IDENTITY) {
            s.gx = gx; s.gy = gy; s.gz = gz;
            s.px = h.px + nd.off[0]; s.py = h.py + nd.off[1]; s.pz = h.pz + nd.off[2];
        } else {
            mulvm(nd.MinvT, gx, gy, gz, s.gx, s.gy, s.gz);   // transform.normal (transform.d:78-81); a positive rescale of g keeps its direction
            double x, y, z;
            mulvm(nd.M, h.px, h.py, h.pz, x, y, z);          // transform.point (transform.d:57-63)
            s.px = x + nd.off[0]; s.py = y + nd.off[1]; s.pz = z + nd.off[2];
        }
    }
    if (nd.kind == KIND_PLANE_W) {   // (0, 1, 0): nothing to normalise
        s.nx = 0.f; s.ny = 1.f; s.nz = 0.f;
        return;
    }
    float fx = (float)s.gx, fy = (float)s.gy, fz = (float)s.gz;
    float inv = rsqrt_fast(dot3f(fx, fy, fz, fx, fy, fz));
    s.nx = fx * inv; s.ny = fy * inv; s.nz = fz * inv;
}

// Per-lane form: the plane-only scene classes (shadow rays among planes are settled by signs, occluded_planes)
template <int MODE>
__device__ __forceinline__ Col shade(const FrameParams& fp, const Ray& ray, const HitRec& h, unsigned& n_shadow) {
    static_assert(plane_only(MODE), "scene classes with bounded / generic nodes shade inside trace_warp");
    constexpr bool SOLO = (MODE & MODE_SOLO) != 0;
    const DevShader& sh = c_scene.shaders[SOLO ? 0 : c_scene.nodes[h.node].shader];
    const bool has_tex = SOLO ? (MODE & MODE_TEX_MASK) != 0 : sh.tex >= 0;
    Surface s;
    surface_of<MODE>(h, &ray, has_tex, s);
    // faceforward (imported_types.d:69-73): the sign decision in FP64, the vector itself in FP32
    float Nx = s.nx, Ny = s.ny, Nz = s.nz;
    // (plane-only scene classes: the geometric normal is (0, 1, 0), the dot product is ray.dy)
    // (a fixed camera off the scene's one plane only hits it with d.y pointing at it: the sign is the camera's side)
    const bool regular = solo_regular<MODE>(fp);
    const bool facing = regular ? fp.solo_side > 0
                        : plane_only(MODE) ? ray.dy < 0 : dot3(ray.dx, ray.dy, ray.dz, s.gx, s.gy, s.gz) < 0;
    if (!facing) { Nx = -Nx; Ny = -Ny; Nz = -Nz; }
    Col diffuse = has_tex ? sample_texture<MODE>(sh.tex, s.u, s.v) : mkcol(sh.color[0], sh.color[1], sh.color[2]);
    Col lightContrib = mkcol(fp.ambient[0], fp.ambient[1], fp.ambient[2]);
    Col specular = mkcol(0.f, 0.f, 0.f);
    const bool phong = SOLO ? (MODE & MODE_PHONG) != 0 : sh.type == C2RT_SHADER_PHONG;
    const int nl = SOLO ? 1 : c_scene.n_lights;
    // shadow-ray origin p + N * 1e-6 (shader.d:88,219)
    // (plane-only scene classes: N = (0, +-1, 0), the x and z terms add an exact zero)
    const double fx = plane_only(MODE) ? s.px : s.px + (double)Nx * 1e-6;
    const double fy = plane_only(MODE) ? s.py + (facing ? 1e-6 : -1e-6) : s.py + (double)Ny * 1e-6;
    const double fz = plane_only(MODE) ? s.pz : s.pz + (double)Nz * 1e-6;
    for (int li = 0; li < nl; li++) {
        const DevLight& L = c_scene.lights[SOLO ? 0 : li];
        // one sample per PointLight (light.d:56-59): avg / numSamples is a division by 1.0f
        if (!L.lit) continue;
        n_shadow++;
        double Dx, Dy, Dz;
        float fDx, fDy, fDz, d2;
        if (plane_only(MODE)) {
            // among planes only the sign of D.y decides visibility: D.y in FP64, the rest of the light vector in FP32
            Dy = L.pos[1] - fy;
            if constexpr (!solo_fast(MODE)) {   // (regular one-plane frames: the plane cannot shadow itself, see isect_plane_solo)
                if (!regular && occluded_planes<MODE>(fx, fy, fz, L, Dy)) continue;
            }
            fDx = L.posf[0] - (float)fx; fDy = (float)Dy; fDz = L.posf[2] - (float)fz;
            d2 = dot3f(fDx, fDy, fDz, fDx, fDy, fDz);
            if (d2 < L.near2) {
                // the hit point is close to the light compared with the light's distance from the origin: the FP32
                // difference above cancels (relative error ~6e-8 (|L| + |p|) / |D|) exactly where 1 / d2 makes the pixel
                // hundreds of times over-bright — take the horizontal part from the FP64 difference there
                fDx = (float)(L.pos[0] - fx); fDz = (float)(L.pos[2] - fz);
                d2 = dot3f(fDx, fDy, fDz, fDx, fDy, fDz);
            }
        } else {
            Dx = Dy = Dz = 0; fDx = fDy = fDz = 0.f; d2 = 1.f;   // (not instantiated: see the static_assert)
        }
        // lighting in FP32 (the reference narrows every factor to float before it touches a Color: SURVEY.md App. C.1)
        const float rs = rsqrt_fast(d2);
        const float lx = fDx * rs, ly = fDy * rs, lz = fDz * rs;
        const float inv_d2 = rs * rs;
        const float cosTheta = plane_only(MODE) ? (facing ? ly : -ly) : dot3f(lx, ly, lz, Nx, Ny, Nz);
        const float br = L.color[0] * inv_d2, bg = L.color[1] * inv_d2, bb = L.color[2] * inv_d2;
        if (cosTheta > 0) {
            lightContrib.r = fmaf(br, cosTheta, lightContrib.r);
            lightContrib.g = fmaf(bg, cosTheta, lightContrib.g);
            lightContrib.b = fmaf(bb, cosTheta, lightContrib.b);
        }
        if (phong) {
            float pw;
            if (sh.exponent <= 2048.0) {
                // reflect(-lightDir, N) . (-ray.dir)  (imported_types.d:62-67, shader.d:235-239)
                const float k = 2.f * cosTheta;
                const float rx = fmaf(k, Nx, -lx), ry = fmaf(k, Ny, -ly), rz = fmaf(k, Nz, -lz);
                float vx = (float)ray.dx, vy = (float)ray.dy, vz = (float)ray.dz;
                if (plane_only(MODE)) {   // un-normalised camera ray (gen_ray)
                    const float iv = rsqrt_fast((float)ray.l2);
                    vx *= iv; vy *= iv; vz *= iv;
                }
                const float cosGamma = -dot3f(rx, ry, rz, vx, vy, vz);
                pw = cosGamma > 0 ? powf(cosGamma, (float)sh.exponent) : 0.f;
            } else {
                // very sharp lobes amplify FP32 rounding of cosGamma by `exponent`: keep FP64 here
                if (plane_only(MODE)) { Dx = L.pos[0] - fx; Dz = L.pos[2] - fz; }
                double ldx = Dx, ldy = Dy, ldz = Dz;
                normalize3(ldx, ldy, ldz);
                double nx = Nx, ny = Ny, nz = Nz;
                normalize3(nx, ny, nz);
                double k = 2 * dot3(ldx, ldy, ldz, nx, ny, nz);
                double rx = k * nx - ldx, ry = k * ny - ldy, rz = k * nz - ldz;
                normalize3(rx, ry, rz);
                double cg = -dot3(rx, ry, rz, ray.dx, ray.dy, ray.dz);
                if (plane_only(MODE)) cg *= rsqrt64(ray.l2);   // un-normalised camera ray (gen_ray)
                pw = cg > 0 ? (float)pow(cg, sh.exponent) : 0.f;
            }
            const float w = pw * sh.strength;
            specular.r = fmaf(br, w, specular.r);
            specular.g = fmaf(bg, w, specular.g);
            specular.b = fmaf(bb, w, specular.b);
        }
    }
    return mkcol(fmaf(diffuse.r, lightContrib.r, specular.r), fmaf(diffuse.g, lightContrib.g, specular.g),
                 fmaf(diffuse.b, lightContrib.b, specular.b));
}

// renderer.d:325-376 (trace) + shader.d:67-105,197-250 (Lambert / Phong with their shadow rays, scene.d:62-78) for the scene
// classes with bounded or generic nodes.  All 32 lanes of the warp call it together; `live` says whether this lane has a ray at
// all, `cam_mask` is the warp's node mask for camera rays.  The camera ray and then one shadow ray per lit light go through the
// SAME node walk (one copy of the cull and of the exact tests in the kernel): phase -1 is the camera ray (closest hit, every
// node of the mask), phase li >= 0 the shadow ray towards light li (any hit; left through __all_sync once every lane that
// needs an answer has its occluder: the reference's early return, scene.d:73-75, taken by the whole warp).
template <int MODE>
__device__ __forceinline__ Col trace_warp(const FrameParams& fp, const Ray& cam_ray, bool live, unsigned& n_shadow, HitRec* out_hit,
                                          const NodeMask& cam_mask) {
    constexpr bool BIG = is_big(MODE);
    __shared__ double s_view[3][BLOCK_THREADS];   // FP64 camera-ray direction, parked for Phong lobes sharper than 2048
    __shared__ NodeWord s_shadow_words[BIG ? WARPS_PER_CTA : 1][BIG ? BIG_MASK_WORDS : 1];   // MODE_BIG: this warp's shadow mask
    NodeMask smask;
    smask.w0 = 0; smask.w1 = 0;
    smask.more = BIG ? s_shadow_words[threadIdx.x >> 5] : nullptr;
    // Register diet (the walk below holds the CSG test's working set; whatever is only needed before or after it is parked in
    // shared memory or recomputed): the FP64 view direction and the surface colour live in s_view / s_diffuse, the shader is
    // kept as an index, the light vector D = L - origin is recomputed where it is needed instead of being carried.
    __shared__ float s_diffuse[3][BLOCK_THREADS];
    Ray r = cam_ray;
    HitRec h;
    h.dist = K_1E99;
    h.node = -1;
    float tmaxf = 3.0e38f;   // (not +inf = (float)1e99: ptxas would prove tmaxf == (float)h.dist * k and re-convert it in every iteration)
    float rsf = 0.f;         // shadow phases: 1 / |D| in FP32
    bool walk = true;        // (a shadow phase in which no lane of the warp has a ray skips the walk)
    bool want = live, found = false, ray_ready = true;
    // shading state of this lane's hit
    bool hit = false;
    int shi = -1;            // shader record of the hit
    float Nx = 0.f, Ny = 0.f, Nz = 0.f;
    Col specular = mkcol(0.f, 0.f, 0.f);
    Col lightContrib = mkcol(fp.ambient[0], fp.ambient[1], fp.ambient[2]);
    double sy = r.dy;        // the ray's vertical direction up to a positive factor (planes are first settled by its sign)
    const int nl = c_scene.n_lit;   // lights with intensity != 0 (shader.d:88,219), listed at scene create
    int li = -1;
    for (;;) {
        const bool anyhit = li >= 0;
        // ---- the node walk: warp-uniform over the mask, in scene order; per-lane work predicated
        const NodeMask& mask = anyhit ? smask : cam_mask;
#pragma unroll 1
        for (int wd = 0; walk && wd < mask_words(); wd++) {
        NodeWord bits = mask_word<BIG>(mask, wd);
#pragma unroll 1
        for (; bits; bits &= bits - 1) {
            const int i = 32 * wd + __ffs((int)bits) - 1;
            const DevNode& nd = node_at<BIG>(i);
            if (want && !found) {
                bool skip = false;
                if (nd.kind == KIND_PLANE_W) {
                    // implied by geometry.d:35-36 (dir.y has the sign of D.y): the common "above the floor, looking / lit from above" case
                    const double y = nd.wp[0];
                    skip = (r.oy > y && sy >= 0) || (r.oy < y && sy <= 0);
                } else if ((MODE & MODE_BOUNDED) && !(nd.flags & NODE_UNBOUNDED)) {
                    skip = cull(nd, r, tmaxf);
                }
                if (!skip) {
                    if (!ray_ready) {   // the FP64 normalisation of a shadow ray (scene.d:66-71), only once a node survives
                        const DevLight& L = c_scene.lights[c_scene.lit[li]];
                        const double Dx = L.pos[0] - r.ox, Dy = L.pos[1] - r.oy, Dz = L.pos[2] - r.oz;
                        const double len2 = dot3(Dx, Dy, Dz, Dx, Dy, Dz);
                        const double inv = rsqrt64(len2);
                        r.dx = Dx * inv; r.dy = Dy * inv; r.dz = Dz * inv;
                        h.dist = len2 * inv;
                        ray_ready = true;
                    }
                    if (node_hit<MODE>(i, nd, r, h)) {
                        if (anyhit) found = true;
                        else if (MODE & MODE_BOUNDED) tmaxf = (float)h.dist * 1.000001f;
                    }
                }
            }
            if (anyhit && __all_sync(FULL_WARP, !want || found)) { walk = false; break; }   // every lane that asked has its occluder
        }
        }
        if (!anyhit) {
            // ---- the camera ray is traced: complete the hit (per lane)
            if (out_hit) *out_hit = h;
            hit = live && h.node >= 0;
            s_view[0][threadIdx.x] = r.dx; s_view[1][threadIdx.x] = r.dy; s_view[2][threadIdx.x] = r.dz;
            if (hit) {
                shi = node_at<BIG>(h.node).shader;
                const DevShader& sh = shader_at<BIG>(shi);
                const bool has_tex = sh.tex >= 0;
                Surface s;
                surface_of<MODE>(h, nullptr, has_tex, s);
                // faceforward (imported_types.d:69-73): the sign decision in FP64, the vector itself in FP32
                Nx = s.nx; Ny = s.ny; Nz = s.nz;
                if (!(dot3(r.dx, r.dy, r.dz, s.gx, s.gy, s.gz) < 0)) { Nx = -Nx; Ny = -Ny; Nz = -Nz; }
                const Col diffuse = has_tex ? sample_texture<MODE>(sh.tex, s.u, s.v) : mkcol(sh.color[0], sh.color[1], sh.color[2]);
                s_diffuse[0][threadIdx.x] = diffuse.r; s_diffuse[1][threadIdx.x] = diffuse.g; s_diffuse[2][threadIdx.x] = diffuse.b;
                // shadow-ray origin p + N * 1e-6 (shader.d:88,219)
                r.ox = s.px + (double)Nx * K_1EM6; r.oy = s.py + (double)Ny * K_1EM6; r.oz = s.pz + (double)Nz * K_1EM6;
            } else if (live) {
                const Col env = env_lookup(r.dx, r.dy, r.dz);   // a miss: renderer.d:366-368 (black unless the cubemap extension is on)
                s_diffuse[0][threadIdx.x] = env.r; s_diffuse[1][threadIdx.x] = env.g; s_diffuse[2][threadIdx.x] = env.b;
            }
        } else if (want && !found) {
            // ---- light li is visible from this lane's hit: lighting in FP32 (the reference narrows every factor to float
            // before it touches a Color: SURVEY.md App. C.1)
            const DevLight& L = c_scene.lights[c_scene.lit[li]];
            float lx, ly, lz, rs;
            if (MODE & MODE_BOUNDED) {   // (float)D * rsqrtf((float)|D|^2): the cull's FP32 direction is exactly that
                lx = r.fdx; ly = r.fdy; lz = r.fdz; rs = rsf;
            } else {
                const double Dx = L.pos[0] - r.ox, Dy = L.pos[1] - r.oy, Dz = L.pos[2] - r.oz;
                rs = rsqrt_fast((float)dot3(Dx, Dy, Dz, Dx, Dy, Dz));
                lx = (float)Dx * rs; ly = (float)Dy * rs; lz = (float)Dz * rs;
            }
            const float inv_d2 = rs * rs;
            const float cosTheta = dot3f(lx, ly, lz, Nx, Ny, Nz);
            const float br = L.color[0] * inv_d2, bg = L.color[1] * inv_d2, bb = L.color[2] * inv_d2;
            if (cosTheta > 0) {
                lightContrib.r = fmaf(br, cosTheta, lightContrib.r);
                lightContrib.g = fmaf(bg, cosTheta, lightContrib.g);
                lightContrib.b = fmaf(bb, cosTheta, lightContrib.b);
            }
            const DevShader& sh = shader_at<BIG>(shi);
            if (sh.type == C2RT_SHADER_PHONG) {
                float pw;
                if (sh.exponent <= 2048.0) {
                    // reflect(-lightDir, N) . (-ray.dir)  (imported_types.d:62-67, shader.d:235-239)
                    const float k = 2.f * cosTheta;
                    const float rx = fmaf(k, Nx, -lx), ry = fmaf(k, Ny, -ly), rz = fmaf(k, Nz, -lz);
                    const float cosGamma = -dot3f(rx, ry, rz, (float)s_view[0][threadIdx.x], (float)s_view[1][threadIdx.x], (float)s_view[2][threadIdx.x]);
                    pw = cosGamma > 0 ? powf(cosGamma, (float)sh.exponent) : 0.f;
                } else {
                    // very sharp lobes amplify FP32 rounding of cosGamma by `exponent`: keep FP64 here
                    double ldx = L.pos[0] - r.ox, ldy = L.pos[1] - r.oy, ldz = L.pos[2] - r.oz;
                    normalize3(ldx, ldy, ldz);
                    double nx = Nx, ny = Ny, nz = Nz;
                    normalize3(nx, ny, nz);
                    double k = 2 * dot3(ldx, ldy, ldz, nx, ny, nz);
                    double rx = k * nx - ldx, ry = k * ny - ldy, rz = k * nz - ldz;
                    normalize3(rx, ry, rz);
                    double cg = -dot3(rx, ry, rz, s_view[0][threadIdx.x], s_view[1][threadIdx.x], s_view[2][threadIdx.x]);
                    pw = cg > 0 ? (float)pow(cg, sh.exponent) : 0.f;
                }
                const float w = pw * sh.strength;
                specular.r = fmaf(br, w, specular.r);
                specular.g = fmaf(bg, w, specular.g);
                specular.b = fmaf(bb, w, specular.b);
            }
        }
        // ---- next lit light (one sample per PointLight, light.d:56-59: avg / numSamples is a division by 1.0f)
        if (++li >= nl) break;
        const DevLight& L = c_scene.lights[c_scene.lit[li]];
        want = hit;
        found = false;
        ray_ready = false;
        n_shadow += hit;
        sy = L.pos[1] - r.oy;
        if (MODE & MODE_BOUNDED) {   // FP32 shadow of the ray for the conservative cull
            const double Dx = L.pos[0] - r.ox, Dz = L.pos[2] - r.oz;
            const float l2f = (float)dot3(Dx, sy, Dz, Dx, sy, Dz);
            rsf = rsqrt_fast(l2f);
            r.fox = cvt_keep(r.ox); r.foy = cvt_keep(r.oy); r.foz = cvt_keep(r.oz);
            r.fdx = cvt_keep(Dx) * rsf; r.fdy = cvt_keep(sy) * rsf; r.fdz = cvt_keep(Dz) * rsf;
            r.olen = sqrt_up(dot3f(r.fox, r.foy, r.foz, r.fox, r.foy, r.foz));   // (only feeds the cull's rounding margins)
            tmaxf = l2f * rsf * 1.000001f;
        }
        shadow_mask<MODE>(smask, fp.bounds, want, r, L, walk);   // walk = false: no lane of this warp has a shadow ray for this light
    }
    const float dr = s_diffuse[0][threadIdx.x], dg = s_diffuse[1][threadIdx.x], db = s_diffuse[2][threadIdx.x];
    if (!hit) return mkcol(dr, dg, db);   // miss: the environment's colour (environment.d:7-10: black), computed in the camera phase
    return mkcol(fmaf(dr, lightContrib.r, specular.r), fmaf(dg, lightContrib.g, specular.g), fmaf(db, lightContrib.b, specular.b));
}

// renderer.d:325-376 for the plane-only scene classes (per lane; only lanes with a ray call it)
template <int MODE>
__device__ __forceinline__ Col trace(const FrameParams& fp, const Ray& ray, unsigned& n_shadow, HitRec* out_hit) {
    HitRec h;
    h.dist = 1e99;
    h.node = -1;
    if constexpr (solo_fast(MODE)) {
        if (isect_plane_solo(fp, fp.solo_h, ray.dx, ray.dy, ray.dz, h.dist)) { h.node = 0; h.leaf = c_scene.nodes[0].geom; h.face = 0; }
    } else if constexpr ((MODE & MODE_SOLO) != 0) {
        if (fp.solo_fast) {
            if (isect_plane_solo(fp, ray.oy - c_scene.nodes[0].wp[0], ray.dx, ray.dy, ray.dz, h.dist)) { h.node = 0; h.leaf = c_scene.nodes[0].geom; h.face = 0; }
        } else node_exact<MODE, true>(0, c_scene.nodes[0], ray, h);
    } else {
#pragma unroll 1
        for (int i = 0; i < c_scene.n_nodes; i++) node_exact<MODE, true>(i, c_scene.nodes[i], ray, h);
    }
    if (out_hit) {
        *out_hit = h;
        if (h.node >= 0) { out_hit->px = fma(ray.dx, h.dist, ray.ox); out_hit->py = fma(ray.dy, h.dist, ray.oy); out_hit->pz = fma(ray.dz, h.dist, ray.oz); }
    }
    if (h.node < 0) return env_lookup(ray.dx, ray.dy, ray.dz);  // renderer.d:366-368
    return shade<MODE>(fp, ray, h, n_shadow);
}
template <int MODE>
__device__ __forceinline__ Col trace_any(const FrameParams& fp, const Ray& ray, bool live, unsigned& n_shadow, HitRec* out_hit, const NodeMask& cam) {
    if constexpr (plane_only(MODE)) return trace<MODE>(fp, ray, n_shadow, out_hit);
    else return trace_warp<MODE>(fp, ray, live, n_shadow, out_hit, cam);
}

// renderer.d:254-313 renderSample (default and DOF branches).  (bx, by, bz) is the un-normalised
// direction through the pixel corner; tap k adds the per-frame constant fp.tap_d[k].
// color.d:10-15 combineStereo + :76-82 adjustSaturation(0.25): anaglyph of the two eye colours
__device__ __forceinline__ Col desaturate(Col c) {
    const float mid = (c.r + c.g + c.b) / 3;
    return mkcol(c.r * 0.25f + mid * (1 - 0.25f), c.g * 0.25f + mid * (1 - 0.25f), c.b * 0.25f + mid * (1 - 0.25f));
}
__device__ __forceinline__ Col combine_stereo(Col left, Col right) {
    left = desaturate(left);
    right = desaturate(right);
    return mkcol(left.r * 1.f + right.r * 0.f, left.g * 0.f + right.g * 1.f, left.b * 0.f + right.b * 1.f);
}

template <int MODE>
__device__ __forceinline__ Col render_sample(const FrameParams& fp, double bx, double by, double bz, double x, double y, uint32_t px,
                                             uint32_t py, uint32_t tap, double jw, double jh, bool live, unsigned& n_primary, unsigned& n_shadow,
                                             HitRec* out_hit, const NodeMask& cam) {
    Ray r;
    if (!(MODE & MODE_SAMPLING)) {               // renderSampleDefault without stereo (renderer.d:303-306): one ray
        uint32_t draw = 0;
        n_primary += live;
        gen_ray<MODE>(fp, bx + fp.tap_d[tap][0], by + fp.tap_d[tap][1], bz + fp.tap_d[tap][2], px, py, tap, 0, draw, 0, r);
        return trace_any<MODE>(fp, r, live, n_shadow, out_hit, cam);
    }
    const bool stereo = fp.stereo_sep != 0;      // renderer.d:276-284,305-312: one ray per eye, then combineStereo
    const int n_eyes = stereo ? 2 : 1;
    const uint32_t n_samples = fp.dof ? fp.num_samples : 1u;
    Col avg = mkcol(0.f, 0.f, 0.f);
#pragma unroll 1
    for (uint32_t i = 0; i < n_samples; i++) {
        uint32_t draw = 0;
        Col left = mkcol(0.f, 0.f, 0.f), c = left;
#pragma unroll 1
        for (int e = 0; e < n_eyes; e++) {
            double vx, vy, vz;
            if (fp.dof) {
                // each call draws its own pixel jitter, then (inside getScreenRay) its own lens sample
                double jx = x + c_tap_x[tap] + uniform01(fp, px, py, tap, i, draw) * jw;   // x + uniform * dx (renderer.d:277)
                double jy = y + c_tap_y[tap] + uniform01(fp, px, py, tap, i, draw) * jh;
                screen_dir(fp, jx, jy, vx, vy, vz);
            } else {
                vx = bx + fp.tap_d[tap][0]; vy = by + fp.tap_d[tap][1]; vz = bz + fp.tap_d[tap][2];
            }
            n_primary += live;
            gen_ray<MODE>(fp, vx, vy, vz, px, py, tap, i, draw, stereo ? (e ? +1 : -1) : 0, r);
            c = trace_any<MODE>(fp, r, live, n_shadow, (out_hit && i == 0 && e == 0) ? out_hit : nullptr, cam);
            if (e == 0) left = c;
        }
        if (stereo) c = combine_stereo(left, c);
        if (!fp.dof) return c;
        avg.r += c.r; avg.g += c.g; avg.b += c.b;
    }
    float n = (float)fp.num_samples;
    return mkcol(avg.r / n, avg.g / n, avg.b / n);
}

__device__ __forceinline__ uint32_t lut8(const uint8_t* lut, float x) {  // color.d:209-214
    if (x <= 0.f) return 0u;
    if (x >= 1.f) return 255u;
    return (uint32_t)__ldg(&lut[(int)(x * 4096.0f)]);
}
__device__ __forceinline__ uint32_t pack_rgb32(const uint8_t* lut, Col c) {  // color.d:154-162
    return lut8(lut, c.b) | (lut8(lut, c.g) << 8) | (lut8(lut, c.r) << 16);
}

// x / 5.f, correctly rounded, for the mean of the five AA taps (renderer.d:249): the fast path of the IEEE division — refined
// reciprocal y, q = x y, one residual correction — written out, without its range check and out-of-line slow path (a colour sum
// is never near the ends of the FP32 range; 0 stays 0).  y = 0.2f + 0.2f (1 - 5 * 0.2f) in FP32 is 0.2f itself.
__device__ __forceinline__ float div5(float x) {
    const float y = fmaf(fmaf(0.2f, -5.f, 1.f), 0.2f, 0.2f);
    const float q = x * y;
    return fmaf(y, fmaf(q, -5.f, x), q);
}

// ---------------------------------------------------------------- per-warp camera-ray node mask
// Which nodes can a camera ray of this warp's 8x4 pixel patch reach?  All those rays start at the camera position and pass
// through the screen rectangle [x0, x0+8] x [y0, y0+4] (pixel corners plus the AA tap offsets <= 0.6), so they lie in the
// circular cone around the normalised sum of the four corner directions whose half-angle reaches the farthest corner (a circular
// cone of less than 90 degrees is convex and contains the convex cone the corners span).  A node whose bounding sphere does not
// touch that cone cannot be hit by any of them.  Lane l tests nodes l, l + 32, ...; one ballot per 32 nodes gives the mask.
// FP64 throughout (once per pixel, amortised over the AA taps), radius padded by 1e-6 (r + distance): the sphere is already
// a conservative bound, this test only has to never drop a reachable node.
constexpr int PATCH_W = 8, PATCH_H = 4;   // pixels of one warp
struct PatchCone {
    double ax, ay, az, cosphi, sinphi;
    bool wide;   // wider than 60 degrees (tiny frames): no culling
};
__device__ __forceinline__ PatchCone patch_cone(const FrameParams& fp, uint32_t x0, uint32_t y0) {
    PatchCone c;
    double u[4][3];
    c.ax = 0; c.ay = 0; c.az = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const double sx = (double)x0 + ((k & 1) ? (double)PATCH_W : -0.01), sy = (double)y0 + ((k & 2) ? (double)PATCH_H : -0.01);
        screen_dir(fp, sx, sy, u[k][0], u[k][1], u[k][2]);
        normalize3(u[k][0], u[k][1], u[k][2]);
        c.ax += u[k][0]; c.ay += u[k][1]; c.az += u[k][2];
    }
    normalize3(c.ax, c.ay, c.az);
    c.cosphi = 1.0;
#pragma unroll
    for (int k = 0; k < 4; k++) c.cosphi = fmin(c.cosphi, dot3(c.ax, c.ay, c.az, u[k][0], u[k][1], u[k][2]));
    c.wide = !(c.cosphi > 0.5);
    c.sinphi = sqrt(fmax(0.0, 1.0 - c.cosphi * c.cosphi));
    return c;
}
__device__ __forceinline__ bool cone_reaches_node(const FrameParams& fp, const PatchCone& c, const float4 b) {
    if (b.w < 0.f || c.wide) return true;   // unbounded node (planes) / a patch too wide to cull
    const double vx = (double)b.x - fp.pos[0], vy = (double)b.y - fp.pos[1], vz = (double)b.z - fp.pos[2];
    const double d2 = dot3(vx, vy, vz, vx, vy, vz);
    const double d = sqrt(d2);
    const double r = (double)b.w * (1.0 + 1e-6) + 1e-6 * d;
    if (!(d > r)) return true;                       // the camera is inside the sphere (or a NaN bound): keep
    const double xa = dot3(vx, vy, vz, c.ax, c.ay, c.az);  // along the axis
    const double ya = sqrt(fmax(0.0, d2 - xa * xa));       // away from it
    if (xa * c.cosphi + ya * c.sinphi >= 0) return !(ya * c.cosphi - xa * c.sinphi > r);   // nearest cone point on the lateral surface
    return false;                                    // nearest cone point is the apex, and d > r
}
template <int MODE>
__device__ __forceinline__ void camera_mask(NodeMask& m, const FrameParams& fp, uint32_t x0, uint32_t y0) {
    constexpr bool BIG = is_big(MODE);
    if (!camera_masked(MODE)) { all_nodes_mask<BIG>(m); return; }
    const PatchCone c = patch_cone(fp, x0, y0);
    ballot_nodes<BIG>(m, [&](int i) { return cone_reaches_node(fp, c, __ldg(&fp.bounds[i])); });   // (global, not constant: per-lane index)
}

// ---------------------------------------------------------------- frame kernel
template <int MODE, int MIN_BLOCKS>
__global__ void __launch_bounds__(BLOCK_THREADS, MIN_BLOCKS) render_frame_kernel(const FrameParams fp) {
    __shared__ __align__(16) float s_rgb[TILE_H][TILE_W * 3];
    __shared__ double s_base[plane_only(MODE) ? 1 : 3][plane_only(MODE) ? 1 : BLOCK_THREADS];

    // tile -> rows: local tile l of this rank belongs to its band (l / tiles_per_band), which is
    // global band (band_local * n_ranks + rank)
    const uint32_t l = blockIdx.y + fp.tile_row0;
    uint32_t tile_row = l;   // one rank owns every band
    if (fp.n_ranks > 1) {
        const uint32_t band_local = l / fp.tiles_per_band;
        const uint32_t within = l - band_local * fp.tiles_per_band;
        tile_row = (band_local * fp.n_ranks + fp.rank) * fp.tiles_per_band + within;
    }
    const uint32_t y0 = tile_row * TILE_H;
    const uint32_t x0 = blockIdx.x * TILE_W;
    // 4 warps, each an 8x4 pixel patch
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t WARPS_X = TILE_W / 8;
    const uint32_t lx = (warp % WARPS_X) * 8 + (lane & 7);
    const uint32_t ly = (warp / WARPS_X) * 4 + (lane >> 3);
    const uint32_t x = x0 + lx, y = y0 + ly;
    const bool active = x < fp.W && y < fp.H;
    if constexpr (solo_bitmap(MODE)) {   // stage the palette (bitmap_fetch); textures without one never read it
        const DevTex& t = c_scene.textures[0];
        if (t.quads) {
            s_solo_palette[threadIdx.x] = __ldg(&t.palette[threadIdx.x]);
            s_solo_palette[threadIdx.x + BLOCK_THREADS] = __ldg(&t.palette[threadIdx.x + BLOCK_THREADS]);
        }
        __syncthreads();
    }
    // c2rt_cancel (renderer.d:93-97,129,147,180: a stop request ends the frame early): a CTA that starts after the flag was
    // raised leaves its tile as it is.  The load is issued here and consumed below, behind the mask / ray set-up.
    int cancelled = 0;
    if (fp.cancel) cancelled = *(const volatile int*)fp.cancel;

    // nodes this warp's camera rays can reach (lane l tests nodes l, l + 32; warp ballot)
    __shared__ NodeWord s_cam_words[is_big(MODE) ? WARPS_PER_CTA : 1][is_big(MODE) ? BIG_MASK_WORDS : 1];   // MODE_BIG: a row per warp
    NodeMask cam;
    cam.w0 = 0; cam.w1 = 0;
    cam.more = is_big(MODE) ? s_cam_words[warp] : nullptr;
    if constexpr (!plane_only(MODE)) camera_mask<MODE>(cam, fp, x0 + (warp % WARPS_X) * PATCH_W, y0 + (warp / WARPS_X) * PATCH_H);

    unsigned n_primary = 0, n_shadow = 0;
    Col c = mkcol(0.f, 0.f, 0.f);
    if (cancelled) return;   // (block-uniform: every thread read the same word)
    if (fp.gi) c = mkcol(fp.gi_fill, fp.gi_fill, fp.gi_fill);   // renderSampleGI: provably black (c2rt_api.cu fill_params)
    else if (active || !plane_only(MODE)) {   // (scene classes with warp masks: every lane runs, lanes off the frame carry no ray)
        // renderer.d:223-251: tap 0 at the pixel corner, then +(.3,.3) (.6,0) (0,.6) (.6,.6); mean of 5 in FP32
        int taps = fp.aa ? 5 : 1;
        uint32_t sx = x, sy = y;
        double jw = 1.0, jh = 1.0;
        if ((MODE & MODE_SAMPLING) && fp.prepass_bucket) {   // preview frames run on the general (sampling) kernels only
            // prepassOnly (renderer.d:110-130): one sample at the corner of the pixel's 16x16 block (blocks are laid out
            // inside each bucket, clipped to it), jitter extent = block size, the colour replicated over the block
            const uint32_t B = fp.prepass_bucket;
            const uint32_t bx0 = x / B * B, by0 = y / B * B;
            const uint32_t rw = min(B, fp.W - min(bx0, fp.W)), rh = min(B, fp.H - min(by0, fp.H));
            const uint32_t dxl = (x - bx0) / 16 * 16, dyl = (y - by0) / 16 * 16;
            sx = bx0 + dxl; sy = by0 + dyl;
            jw = (double)(min(rw, dxl + 16) - dxl);
            jh = (double)(min(rh, dyl + 16) - dyl);
            taps = 1;
        }
        const double xd = (double)sx, yd = (double)sy;
        double bx, by, bz;
        screen_dir(fp, xd, yd, bx, by, bz);
        if constexpr (!plane_only(MODE)) {   // the pixel's base direction waits in shared memory while a tap is traced (register diet)
            s_base[0][threadIdx.x] = bx; s_base[1][threadIdx.x] = by; s_base[2][threadIdx.x] = bz;
        }
#pragma unroll 1
        for (int s = 0; s < taps; s++) {
            if constexpr (!plane_only(MODE)) {
                bx = ((volatile double*)s_base[0])[threadIdx.x]; by = ((volatile double*)s_base[1])[threadIdx.x]; bz = ((volatile double*)s_base[2])[threadIdx.x];
            }
            Col t = render_sample<MODE>(fp, bx, by, bz, xd, yd, sx, sy, s, jw, jh, active, n_primary, n_shadow, nullptr, cam);
            c.r += t.r; c.g += t.g; c.b += t.b;
        }
        if (taps == 5) { c.r = div5(c.r); c.g = div5(c.g); c.b = div5(c.b); }  // accum / 5 (renderer.d:249)
    }

    // output row of this tile row: full frame or compact (only this rank's rows, in order)
    const uint32_t out_y0 = fp.compact ? l * TILE_H : y0;
    const bool full_tile = (x0 + TILE_W <= fp.W) && ((fp.W & 3u) == 0);
    if (!fp.rgb) {
        // ARGB-only delivery (an interactive host that blits the packed plane): no float frame is written
    } else if (full_tile) {
        s_rgb[ly][lx * 3 + 0] = c.r;
        s_rgb[ly][lx * 3 + 1] = c.g;
        s_rgb[ly][lx * 3 + 2] = c.b;
        __syncthreads();
        // 8 rows x 12 float4 = 96 vector stores, 192 contiguous bytes per row
        if (threadIdx.x < TILE_H * (TILE_W * 3 / 4)) {
            const uint32_t row = threadIdx.x / (TILE_W * 3 / 4), q = threadIdx.x % (TILE_W * 3 / 4);
            if (y0 + row < fp.H) {
                float4 v4 = *reinterpret_cast<const float4*>(&s_rgb[row][q * 4]);
                float* dst = fp.rgb + ((size_t)(out_y0 + row) * fp.W + x0) * 3 + q * 4;
                *reinterpret_cast<float4*>(dst) = v4;
            }
        }
    } else if (active) {
        float* dst = fp.rgb + ((size_t)(out_y0 + ly) * fp.W + x) * 3;
        dst[0] = c.r; dst[1] = c.g; dst[2] = c.b;
    }
    if (fp.argb && active) fp.argb[(size_t)(out_y0 + ly) * fp.W + x] = pack_rgb32(fp.lut, c);

    // Frame-complete wait folded into RANK 0's kernel (c2rt.h c2rt_band.done_flags; one process per GPU, peers store their bands
    // into rank 0's frame over NVLink and raise flags[rank] from a one-thread kernel behind their render kernel — a kernel
    // boundary, so every band store is performed at system scope before the flag).  The last CTA of rank 0's launch to finish
    // waits for all peers' flags: the end of rank 0's kernel IS the complete frame, with no extra launch on its critical path.
    // (A per-CTA __threadfence_system() + flag store inside the PEERS' kernels was measured too: the fence waits for the
    // CTA's remote stores and held every CTA ~4 us — 21 % on the 8K chessboard; profiles/r2_inkernel_fence.log.)
    if (fp.done_flags && threadIdx.x == 0) {
        if (atomicAdd(fp.done_counter, 1u) == gridDim.x * gridDim.y - 1u) {
            *fp.done_counter = 0u;   // ready for the next launch on this device (stream order)
            const long long t0 = clock64();
            for (uint32_t r = 1; r < fp.n_ranks; r++)
                // frame numbers only grow; the signed difference tolerates wrap-around
                while ((int)(ld_acquire_sys(fp.done_flags + r) - fp.frame_no) < 0)
                    if (clock64() - t0 > 4000000000ll) {   // ~2 s: a peer died; do not hang the device, report it
                        atomicAdd(fp.done_flags + fp.n_ranks, 1u);
                        break;
                    }
        }
    }

    if (fp.count_rays) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n_primary += __shfl_xor_sync(0xffffffffu, n_primary, o);
            n_shadow += __shfl_xor_sync(0xffffffffu, n_shadow, o);
        }
        if (lane == 0) {
            atomicAdd(&fp.counters[0], (unsigned long long)n_primary);
            atomicAdd(&fp.counters[1], (unsigned long long)n_shadow);
        }
    }
}

// renderer.d:46-57 renderPixel: one corner sample + the hit record, by a single thread
struct PixelOut {
    float rgb[3];
    int node;
    double dist, p[3], n[3], u, v;
};

template <bool BIG>
__global__ void render_pixel_kernel(const FrameParams fp, int x, int y, PixelOut* out) {
    // one warp; lane 0 carries the ray, the other lanes only take part in the warp-wide votes of trace_warp
    const bool live = threadIdx.x == 0;
    unsigned a = 0, b = 0;
    HitRec h;
    h.node = -1;
    h.dist = 1e99;
    double bx, by, bz;
    screen_dir(fp, (double)x, (double)y, bx, by, bz);
    constexpr int M = MODE_BOUNDED | MODE_GENERIC | MODE_NESTED | MODE_SAMPLING | (BIG ? MODE_BIG : 0);
    __shared__ NodeWord s_cam_words[BIG ? BIG_MASK_WORDS : 1];
    NodeMask cam;
    cam.w0 = 0; cam.w1 = 0;
    cam.more = BIG ? s_cam_words : nullptr;
    all_nodes_mask<BIG>(cam);
    Col c = render_sample<M>(fp, bx, by, bz, (double)x, (double)y, (uint32_t)x, (uint32_t)y, 0, 1.0, 1.0, live, a, b, &h, cam);
    if (!live) return;
    out->rgb[0] = c.r; out->rgb[1] = c.g; out->rgb[2] = c.b;
    out->node = h.node;
    out->dist = h.dist;
    if (h.node >= 0) {
        Surface w;
        surface_of<MODE_BOUNDED | MODE_GENERIC | (BIG ? MODE_BIG : 0)>(h, nullptr, true, w);
        out->p[0] = w.px; out->p[1] = w.py; out->p[2] = w.pz;
        double gx = w.gx, gy = w.gy, gz = w.gz;
        normalize3(gx, gy, gz);
        out->n[0] = gx; out->n[1] = gy; out->n[2] = gz;
        out->u = w.u; out->v = w.v;
    }
}

// rank 0: scatter rank-major compact band buffers into the full frame (see c2rt.h c2rt_deinterleave)
__global__ void deinterleave_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, uint32_t row_words,
                                    uint32_t height, uint32_t n_ranks, uint32_t band_rows, uint32_t rows_pad) {
    const uint32_t y = blockIdx.y;
    if (y >= height) return;
    const uint32_t band = y / band_rows, rank = band % n_ranks, band_local = band / n_ranks;
    const uint32_t local_row = band_local * band_rows + (y - band * band_rows);
    const uint32_t* s = src + ((size_t)rank * rows_pad + local_row) * row_words;
    uint32_t* d = dst + (size_t)y * row_words;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < row_words; i += gridDim.x * blockDim.x) d[i] = s[i];
}

// Device-side start gate between ranks (c2rt.h c2rt_gate): every rank adds 1 to a counter in rank 0's memory and waits until
// all n_ranks of this round have arrived.  Enqueued in front of a frame it (i) starts the ranks' kernels together without a host
// round trip and (ii) keeps a fast peer from storing bands of frame k + 1 while rank 0's stream still works on frame k.
// peer ranks: raise flags[rank] = frame_no behind the render kernel (stream order: the kernel boundary has performed its band
// stores at system scope; the release store orders the flag after them for rank 0's acquire loads)
__global__ void signal_kernel(uint32_t* flag, uint32_t frame_no) {
    __threadfence_system();
    st_release_sys(flag, frame_no);
}
__global__ void gate_kernel(uint32_t* gate, uint32_t round, uint32_t n_ranks, uint32_t* err) {
    const uint32_t target = round * n_ranks;
    __threadfence_system();
    atomicAdd_system(gate, 1u);
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(gate) - target) < 0)
        if (clock64() - t0 > 4000000000ll) {   // ~2 s: a rank never arrived
            atomicAdd_system(err, 1u);
            break;
        }
}

// dependent-free FMA streams for the roofline denominators (c2rt.h c2rt_measure_fma_peak)
template <typename T>
__global__ void fma_peak_kernel(T* out, int iters, T a, T b) {
    T x0 = (T)threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
            x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
        }
    }
    T s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == (T)123456789) out[0] = s;
}

// ---------------------------------------------------------------- host-side launchers (used by c2rt_api.cu)
cudaError_t upload_scene(const DevScene& s, cudaStream_t st) {
    return cudaMemcpyToSymbolAsync(c_scene, &s, sizeof(DevScene), 0, cudaMemcpyHostToDevice, st);
}

#ifndef C2RT_MINBLOCKS_SIMPLE
#define C2RT_MINBLOCKS_SIMPLE 6
#endif
#ifndef C2RT_MINBLOCKS_BOUNDED
#define C2RT_MINBLOCKS_BOUNDED 4
#endif
#ifndef C2RT_MINBLOCKS_SAMPLING
#define C2RT_MINBLOCKS_SAMPLING 4
#endif
#ifndef C2RT_MINBLOCKS_FULL
#define C2RT_MINBLOCKS_FULL 4   // 128 registers: the shared node walk of trace_warp spills at 96 (profiles/r2_variant_sweeps.log)
#endif
#ifndef C2RT_MINBLOCKS_SOLO
#define C2RT_MINBLOCKS_SOLO 7
#endif
#ifndef C2RT_MINBLOCKS_NESTED
#define C2RT_MINBLOCKS_NESTED 2
#endif
#ifndef C2RT_MINBLOCKS_SOLO_SAMPLING
#define C2RT_MINBLOCKS_SOLO_SAMPLING 7   // 71 registers, no spills: C3 14.30 -> 13.74 ms (5: 15.37, 6: 14.30, 8: 13.88; profiles/r2_s2_experiments.log)
#endif

// MODE_SOLO scene classes: texture kind x shader kind, each with and without the DOF / stereo sampling loop
template <int TEXK, int PH>
static void launch_solo(const FrameParams& fp, bool sampling, dim3 grid, cudaStream_t st) {
    constexpr int M = MODE_SOLO | (TEXK << MODE_TEX_SHIFT) | (PH ? MODE_PHONG : 0);
    if (sampling) render_frame_kernel<M | MODE_SAMPLING, C2RT_MINBLOCKS_SOLO_SAMPLING><<<grid, BLOCK_THREADS, 0, st>>>(fp);
    else render_frame_kernel<M, C2RT_MINBLOCKS_SOLO><<<grid, BLOCK_THREADS, 0, st>>>(fp);
}

cudaError_t launch_frame(const FrameParams& fp, int mode, uint32_t local_tile_rows, cudaStream_t st) {
    if (local_tile_rows == 0) return cudaSuccess;
    dim3 grid((fp.W + TILE_W - 1) / TILE_W, local_tile_rows);
    constexpr int FULL = MODE_BOUNDED | MODE_GENERIC, ALL = FULL | MODE_NESTED;
    // (one-plane scenes: a frame whose geometry is not regular — fill_params — runs on the general sampling kernel too)
    const bool sampling = fp.dof || fp.stereo_sep != 0 || fp.prepass_bucket || ((mode & MODE_SOLO) && !fp.solo_fast);
    if (mode & MODE_SOLO) {
        switch (((mode & MODE_TEX_MASK) >> MODE_TEX_SHIFT) | ((mode & MODE_PHONG) ? 4 : 0)) {
            case 0: launch_solo<0, 0>(fp, sampling, grid, st); break;
            case 1: launch_solo<1, 0>(fp, sampling, grid, st); break;
            case 2: launch_solo<2, 0>(fp, sampling, grid, st); break;
            case 3: launch_solo<3, 0>(fp, sampling, grid, st); break;
            case 4: launch_solo<0, 1>(fp, sampling, grid, st); break;
            case 5: launch_solo<1, 1>(fp, sampling, grid, st); break;
            case 6: launch_solo<2, 1>(fp, sampling, grid, st); break;
            default: launch_solo<3, 1>(fp, sampling, grid, st); break;
        }
    } else if (mode & MODE_BIG) {
        // scenes beyond the constant block (records in global memory, multi-word node masks): the general kernels only
        if (sampling) render_frame_kernel<ALL | MODE_BIG | MODE_SAMPLING, C2RT_MINBLOCKS_NESTED><<<grid, BLOCK_THREADS, 0, st>>>(fp);
        else if (mode & MODE_NESTED) render_frame_kernel<ALL | MODE_BIG, C2RT_MINBLOCKS_NESTED><<<grid, BLOCK_THREADS, 0, st>>>(fp);
        else render_frame_kernel<FULL | MODE_BIG, C2RT_MINBLOCKS_FULL><<<grid, BLOCK_THREADS, 0, st>>>(fp);
    } else if (sampling) {
        // DOF / stereo / prepass-only frames: two general kernels only (the per-sample loop dominates, the scene class matters less)
        if (mode & MODE_NESTED) render_frame_kernel<ALL | MODE_SAMPLING, C2RT_MINBLOCKS_NESTED><<<grid, BLOCK_THREADS, 0, st>>>(fp);
        else render_frame_kernel<FULL | MODE_SAMPLING, C2RT_MINBLOCKS_SAMPLING><<<grid, BLOCK_THREADS, 0, st>>>(fp);
    } else if (mode & MODE_NESTED) render_frame_kernel<ALL, C2RT_MINBLOCKS_NESTED><<<grid, BLOCK_THREADS, 0, st>>>(fp);
    else if (mode & MODE_GENERIC) render_frame_kernel<FULL, C2RT_MINBLOCKS_FULL><<<grid, BLOCK_THREADS, 0, st>>>(fp);
    else if (mode & MODE_BOUNDED) render_frame_kernel<MODE_BOUNDED, C2RT_MINBLOCKS_BOUNDED><<<grid, BLOCK_THREADS, 0, st>>>(fp);
    else render_frame_kernel<0, C2RT_MINBLOCKS_SIMPLE><<<grid, BLOCK_THREADS, 0, st>>>(fp);
    return cudaGetLastError();
}

cudaError_t launch_pixel(const FrameParams& fp, int mode, int x, int y, void* d_out, cudaStream_t st) {
    if (mode & MODE_BIG) render_pixel_kernel<true><<<1, 32, 0, st>>>(fp, x, y, (PixelOut*)d_out);
    else render_pixel_kernel<false><<<1, 32, 0, st>>>(fp, x, y, (PixelOut*)d_out);
    return cudaGetLastError();
}

cudaError_t launch_deinterleave(const void* src, void* dst, uint32_t row_words, uint32_t height, uint32_t n_ranks,
                                uint32_t band_rows, uint32_t rows_pad, cudaStream_t st) {
    dim3 grid((row_words + 1023) / 1024 < 1 ? 1 : (row_words + 1023) / 1024, height);
    deinterleave_kernel<<<grid, 256, 0, st>>>((const uint32_t*)src, (uint32_t*)dst, row_words, height, n_ranks, band_rows, rows_pad);
    return cudaGetLastError();
}

cudaError_t launch_signal(void* flag, uint32_t frame_no, cudaStream_t st) {
    signal_kernel<<<1, 1, 0, st>>>((uint32_t*)flag, frame_no);
    return cudaGetLastError();
}
cudaError_t launch_gate(void* gate, uint32_t round, uint32_t n_ranks, void* err, cudaStream_t st) {
    gate_kernel<<<1, 1, 0, st>>>((uint32_t*)gate, round, n_ranks, (uint32_t*)err);
    return cudaGetLastError();
}

cudaError_t launch_fma_peak(bool fp64, int blocks, int threads, int iters, void* d_out, cudaStream_t st) {
    if (fp64) fma_peak_kernel<double><<<blocks, threads, 0, st>>>((double*)d_out, iters, 1.0000001, 1e-9);
    else fma_peak_kernel<float><<<blocks, threads, 0, st>>>((float*)d_out, iters, 1.0000001f, 1e-9f);
    return cudaGetLastError();
}

size_t pixel_out_size() { return sizeof(PixelOut); }

}  // namespace c2rt
