/* c2rt.h — C ABI of libc2rt.so, the B200 (sm_100a) implementation of Chess2RT's per-pixel
 * render loop.  Plain C, no exceptions, no torch/C++ types in any signature.
 *
 * This is the drop-in boundary for the reference's renderer seam:
 *   /root/reference/source/rt/renderer.d:23-44   renderSceneAsync(Scene, Image!Color, isRendering*, needsRendering*)
 *   /root/reference/source/rt/renderer.d:46-57   renderPixel(Scene, Image!Color, x, y) -> (Color, TraceResult)
 *   /root/reference/source/rt/renderer.d:72-189  Renderer(scene, output).renderRT()
 * The D host keeps its loaders and object model; a scene flattener lowers `Scene`
 * (/root/reference/source/rt/scene.d:38-51) into the structure-of-arrays description below
 * (field map: SURVEY.md Appendix B; D-side binding: INTEGRATION.md), and the calls below replace
 * the body of renderRT / renderPixel.  The C++ mirror of that host side lives in
 * chess2rt_b200/host/ (this image has no D toolchain).
 *
 * Conventions
 *   - every function returns C2RT_OK (0) or a negative c2rt_status; c2rt_last_error() gives the
 *     message for the calling thread.  Nothing unwinds across this boundary.
 *   - all pointers in a description are borrowed for the duration of the call only; the library
 *     deep-copies what it keeps.  Output buffers are caller-owned and never retained.
 *   - geometry quantities are FP64, colour quantities FP32, as in the reference
 *     (imported_types.d:10-11 `Vector = vec3d`; color.d:27-35 `Color{float r,g,b}`).
 *   - matrices are 3x3 row-major `m[3*row+col]`, used as row-vector x matrix
 *     (imported_types.d:13-20 `mul`).
 *   - there is no CPU fallback: every entry point that renders fails with C2RT_ERR_CUDA when no
 *     sm_100-class device / driver is usable.
 */
#ifndef C2RT_H
#define C2RT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define C2RT_ABI_VERSION 2

typedef enum c2rt_status {
    C2RT_CANCELLED = 1,          /* not an error: c2rt_render ended early because c2rt_cancel was called (the frame is partial) */
    C2RT_OK = 0,
    C2RT_ERR_INVALID_ARG = -1,   /* null pointer, bad index, inconsistent sizes */
    C2RT_ERR_UNSUPPORTED = -2,   /* input the path cannot honour (GI on a Phong-shaded scene, CSG nesting beyond the limit) */
    C2RT_ERR_CUDA = -3,          /* CUDA runtime error or no usable device */
    C2RT_ERR_NOT_INITIALISED = -4,
    C2RT_ERR_LIMIT = -5          /* scene exceeds a compiled-in capacity (C2RT_MAX_*) */
} c2rt_status;

/* Capacities of the on-chip (constant memory) scene block.  A scene within them renders from constant memory; a larger one
 * (the reference's scene.nodes is unbounded: scene.d:38-51, renderer.d:336-338) keeps its node / geometry / shader / texture
 * records in global memory, up to the *_GLOBAL limits below; beyond those c2rt_scene_create returns C2RT_ERR_LIMIT. */
#define C2RT_MAX_NODES_GLOBAL 4096
#define C2RT_MAX_GEOMS_GLOBAL 16384
#define C2RT_MAX_SHADERS_GLOBAL 4096
#define C2RT_MAX_TEXTURES_GLOBAL 1024
#define C2RT_MAX_NODES 64
#define C2RT_MAX_GEOMS 128
#define C2RT_MAX_SHADERS 64
#define C2RT_MAX_TEXTURES 32
#define C2RT_MAX_LIGHTS 8
#define C2RT_MAX_GPUS 8

/* Geometry kinds — geometry.d:15 Plane, :73 Sphere, :149 Cube, :357 CsgUnion, :367 CsgInter, :377 CsgDiff */
enum { C2RT_GEOM_PLANE = 0, C2RT_GEOM_SPHERE = 1, C2RT_GEOM_CUBE = 2, C2RT_GEOM_CSG_UNION = 3, C2RT_GEOM_CSG_INTER = 4, C2RT_GEOM_CSG_DIFF = 5 };
/* Shader kinds — shader.d:54 Lambert, :177 Phong */
enum { C2RT_SHADER_LAMBERT = 0, C2RT_SHADER_PHONG = 1 };
/* Texture kinds — texture.d:20 Checker, :70 Procedure2, :103 BitmapTexture */
enum { C2RT_TEX_CHECKER = 0, C2RT_TEX_PROCEDURE2 = 1, C2RT_TEX_BITMAP = 2 };
/* Environment kinds — environment.d:5-15: the reference has the black stub only.  CUBEMAP is an EXTENSION (no counterpart in the
 * reference, parity unpinned; DESIGN.md "Cubemap environment"): six bitmap faces, looked up by the ray direction of a miss. */
enum { C2RT_ENV_BLACK = 0, C2RT_ENV_CUBEMAP = 1 };

/* Flattened scene: structure of arrays, indices instead of object references.
 * One D object -> one index; objects shared by several nodes keep one entry (lecture5.sdl's
 * sphere "S" is used by three nodes).  CSG children are geometry indices that must be smaller
 * than the CSG's own index (the loader only resolves names defined above, geometry.d:343-347),
 * so the `current.g is left` identity test (geometry.d:314) becomes an index comparison. */
typedef struct c2rt_scene_desc {
    uint32_t struct_size;          /* sizeof(c2rt_scene_desc), for ABI checking */
    uint32_t abi_version;          /* C2RT_ABI_VERSION */

    /* nodes — node.d:7-10 + transform.d:11-14 */
    uint32_t n_nodes;
    const int32_t* node_geom;      /* [n_nodes] geometry index */
    const int32_t* node_shader;    /* [n_nodes] shader index */
    const double* node_transform;  /* [n_nodes*9] Transform.transform */
    const double* node_inverse;    /* [n_nodes*9] Transform.inverseTransform */
    const double* node_inverse_t;  /* [n_nodes*9] Transform.transposedInverse */
    const double* node_offset;     /* [n_nodes*3] Transform.offset */

    /* geometries — geometry.d */
    uint32_t n_geoms;
    const int32_t* geom_type;      /* [n_geoms] C2RT_GEOM_* */
    const double* geom_params;     /* [n_geoms*4] plane: y, limit (NaN = unbounded, geometry.d:19,46), -, -
                                                   sphere: center xyz, R;  cube: center xyz, side;  csg: unused */
    const int32_t* geom_left;      /* [n_geoms] CSG left child index, -1 otherwise */
    const int32_t* geom_right;     /* [n_geoms] CSG right child index, -1 otherwise */

    /* shaders — shader.d:26,57,179-181 */
    uint32_t n_shaders;
    const int32_t* shader_type;    /* [n_shaders] C2RT_SHADER_* */
    const float* shader_color;     /* [n_shaders*3] */
    const int32_t* shader_texture; /* [n_shaders] texture index or -1 */
    const double* shader_exponent; /* [n_shaders] Phong.exponent (already clamped, shader.d:268) */
    const float* shader_strength;  /* [n_shaders] Phong.strength (already clamped, shader.d:271) */

    /* textures — texture.d:22-23,72-73,151-161 */
    uint32_t n_textures;
    const int32_t* tex_type;       /* [n_textures] C2RT_TEX_* */
    const float* tex_colors;       /* [n_textures*18] checker: color1[3], color2[3];
                                                      procedure2: colorU[3][3] then colorV[3][3]; bitmap: unused */
    const double* tex_params;      /* [n_textures*6]  checker: size; procedure2: freqU[3], freqV[3];
                                                      bitmap: scaling (the D field is a float; widen it exactly) */
    const int32_t* tex_width;      /* [n_textures] bitmap width, 0 otherwise */
    const int32_t* tex_height;     /* [n_textures] bitmap height, 0 otherwise */
    const uint64_t* tex_texel_offset; /* [n_textures] offset, in texels, of this bitmap inside `texels` */
    const float* texels;           /* all bitmaps back to back: r,g,b per texel, row-major, row 0 first —
                                      exactly Image!Color.pixels after the load-time gamma pass
                                      (imageio/image.d:18-54, texture.d:137-141) */
    uint64_t n_texels;

    /* lights — light.d:8-9,54 (PointLight only) */
    uint32_t n_lights;
    const double* light_pos;       /* [n_lights*3] */
    const float* light_color;      /* [n_lights*3] lightColor (NOT premultiplied) */
    const float* light_power;      /* [n_lights]   lightPower */

    /* environment — environment.d:5-15 (what `scene.environment.getEnvironment(ray.dir)` returns for a miss, renderer.d:366-368) */
    int32_t env_type;              /* C2RT_ENV_* */
    int32_t env_reserved;
    int32_t env_face_width[6];     /* CUBEMAP: faces in the order +x, -x, +y, -y, +z, -z */
    int32_t env_face_height[6];
    uint64_t env_face_texel_offset[6]; /* offset, in texels, of each face inside `texels` (post-gamma values like the bitmaps) */
} c2rt_scene_desc;

/* Camera state AFTER Camera.beginFrame (camera.d:77-117) and setFrameSize (camera.d:231-236). */
typedef struct c2rt_camera {
    double pos[3];
    double up_left[3], up_right[3], down_left[3];   /* camera.d:51, world space (pos already added) */
    double right_dir[3], up_dir[3], front_dir[3];   /* camera.d:52 */
    uint32_t frame_width, frame_height;             /* camera.d:29-30: the divisors in getScreenRay */
    int32_t dof;                                    /* camera.d:43 */
    uint32_t num_samples;                           /* camera.d:44 */
    double focal_plane_dist;                        /* camera.d:40 */
    double disc_multiplier;                         /* camera.d:42,252 = 10 / fNumber */
    double stereo_separation;                       /* camera.d:45; 0 = off, otherwise every sample traces a left and a right
                                                       eye ray and combines them (combineStereo, color.d:10-15) */
} c2rt_camera;

/* GlobalSettings fields the render path reads (global_settings.d:8-35) + the pinned-RNG seed. */
typedef struct c2rt_settings {
    uint32_t frame_width, frame_height;   /* output size */
    int32_t aa_enabled;                   /* AAEnabled: 5 samples per pixel for ALL pixels (renderer.d:183-186) */
    int32_t gi_enabled;                   /* GIEnabled: path tracing when the camera has no DOF (renderer.d:256-263).  With the only light
                                             type of the reference (PointLight, solidAngle 0: light.d:72-75) every path returns exactly
                                             black, so the frame is black; a Phong-shaded node makes the reference halt (shader.d:252-262)
                                             and is refused with C2RT_ERR_UNSUPPORTED */
    int32_t prepass_enabled;              /* accepted and ignored: the prepass is fully overwritten (renderer.d:110-142) */
    int32_t prepass_only;                 /* prepassOnly: with prepass_enabled, the frame is the 16x16-block preview of renderer.d:110-130
                                             (one sample per block, replicated); without it nothing is rendered, like the reference */
    uint32_t max_trace_depth;             /* renderer.d:330 (primary rays have depth 0) */
    float ambient_light[3];               /* ambientLightColor */
    uint64_t rng_seed;                    /* DOF only: seed of the pinned counter-based generator (c2rt_rng_u31) */
    int32_t count_rays;                   /* non-zero: fill c2rt_stats.primary_rays / shadow_rays (slightly slower) */
    uint32_t bucket_size;                 /* bucketSize (global_settings.d:16); only the prepass block grid depends on it; 0 = 48 */
    uint32_t paths_per_pixel;             /* pathsPerPixel (renderer.d:293-300): GI frames only; 0 makes the mean 0/0 = NaN like the reference */
    uint32_t reserved;
} c2rt_settings;

/* Interleaved row bands (multi-GPU): row y belongs to rank ((y / band_rows) % n_ranks).
 * band_rows must be a multiple of 8 (the CUDA block tile height). */
typedef struct c2rt_band {
    uint32_t rank, n_ranks, band_rows;
    uint32_t compact;   /* 0: outputs are full frames (row y at y*W); 1: outputs hold only this rank's rows, in order */
    /* Frame-complete signalling for one process per GPU (bands stored into rank 0's frame through a c2rt_frame_import
     * mapping).  NULL: off.  Otherwise uint32 flags[n_ranks + 1] in RANK 0's frame allocation, zero at start: [0] is the
     * start gate's counter (c2rt_gate), [r] the last frame_no rank r > 0 completed, [n_ranks] counts time-outs.  A peer's
     * c2rt_render_device enqueues a one-thread kernel behind its render kernel that stores frame_no into flags[rank]
     * (st.release.sys; the kernel boundary has performed the band stores at system scope).  The last CTA of RANK 0's render
     * kernel waits for flags[1..n_ranks-1] >= frame_no, so rank 0's kernel ends when the whole frame is in its memory — no
     * wait launch on its critical path.  The wait gives up after ~2 s and bumps flags[n_ranks] instead of hanging the device:
     * read that word back before trusting a frame.  frame_no must grow by one per frame (1, 2, ...); one such launch in
     * flight per device. */
    void* done_flags;
    uint32_t frame_no;
    uint32_t reserved;
} c2rt_band;

typedef struct c2rt_stats {
    double kernel_ms;        /* device time of the render kernel(s), CUDA events (max over devices) */
    double total_ms;         /* host wall time of the call, copies included */
    uint64_t primary_rays;   /* only when settings.count_rays */
    uint64_t shadow_rays;
    uint32_t n_gpus;
    uint32_t launches;       /* kernels launched by the call */
} c2rt_stats;

/* renderer.d:14-21 TraceResult, the part a caller can use */
typedef struct c2rt_hit {
    int32_t node;            /* index of closestNode, -1 = miss */
    int32_t reserved;
    double dist;
    double p[3];
    double normal[3];
    double u, v;
} c2rt_hit;

typedef struct c2rt_scene c2rt_scene;

/* Library lifetime.  device_ids == NULL -> devices 0..n_gpus-1.  With n_gpus > 1 c2rt_render splits
 * the frame into interleaved row bands; every device copies its bands to the caller's host frame over
 * its own PCIe link (C2RT_GATHER=root: the frame lives on device_ids[0] and the other devices store
 * their bands straight into it through peer-mapped pointers, NVLink P2P).  The library then owns
 * n_gpus - 1 helper threads (one per extra device, idle between frames; joined by c2rt_shutdown or the
 * next c2rt_init).  Calling c2rt_init again re-initialises. */
int c2rt_init(int n_gpus, const int* device_ids);
void c2rt_shutdown(void);
int c2rt_abi_version(void);
int c2rt_device_count(void);
const char* c2rt_last_error(void);

/* Scene upload: validates, deep-copies, builds the device scene block + bounding volumes. */
int c2rt_scene_create(const c2rt_scene_desc* desc, c2rt_scene** out);
void c2rt_scene_destroy(c2rt_scene* scene);

/* Replaces Renderer.renderRT (renderer.d:83-189): blocking, one frame into caller-owned HOST
 * memory.  rgb: frame_width*frame_height*3 floats laid out as Image!Color.pixels
 * (row-major, top row first, imageio/image.d:45-54).  argb (nullable): one uint32 per pixel,
 * Color.toRGB32 packing (color.d:154-162: r<<16 | g<<8 | b through the sRGB table). */
int c2rt_render(c2rt_scene* scene, const c2rt_camera* camera, const c2rt_settings* settings,
                float* rgb, uint32_t* argb, c2rt_stats* stats);
/* rgb may be NULL when argb is not: ARGB-only delivery for an interactive host that only blits the packed plane
 * (sdl2_gui.d:139-155) — no float frame is written or copied, a 1080p frame is 8.3 MB over PCIe instead of 33 MB. */

/* Replaces the stop request the reference polls between its passes (renderer.d:93-97,129,147,180): callable from ANY thread
 * while another thread is inside c2rt_render.  Raises a flag on every device; tiles whose CTA has not started yet are skipped,
 * tiles in flight finish, and that c2rt_render returns C2RT_CANCELLED with a partially rendered frame (a frame the GUI is about
 * to replace anyway: raytracer_demo.d:102-124).  A cancel that arrives while no frame is in progress cancels nothing. */
int c2rt_cancel(void);

/* Same frame, outputs in DEVICE memory of the CURRENT device, launched on `stream`
 * (a cudaStream_t, NULL = default stream) and NOT synchronised.  One process per GPU uses this
 * with its own band; a single-GPU caller passes band == NULL.  d_argb may be NULL.
 * stats (nullable) gets launches only; timing belongs to the caller's events.
 * d_rgb must be 16-byte aligned when frame_width is a multiple of 4 (bands are written with float4 stores).
 * A device holds ONE scene block at a time: rendering a different scene than the previous call on this device first waits
 * for all work on the device and re-uploads the block, so (i) frames of two scenes never overlap on one device and
 * (ii) the first frame of a scene on a device must be issued outside any CUDA stream capture (later ones may be captured). */
int c2rt_render_device(c2rt_scene* scene, const c2rt_camera* camera, const c2rt_settings* settings,
                       const c2rt_band* band, float* d_rgb, uint32_t* d_argb, void* stream, c2rt_stats* stats);

/* Ray counters accumulated by c2rt_render_device calls with settings.count_rays since the last
 * read on the current device (synchronises `stream`), then reset. */
int c2rt_read_ray_counters(c2rt_scene* scene, void* stream, uint64_t* primary, uint64_t* shadow);

/* Rank 0 side of the band gather: `gathered` holds n_ranks compact band buffers back to back
 * (rank-major, each padded to `rows_per_rank_padded` rows); scatter them into the full frame.
 * elem_floats = 3 for RGB float frames, 1 for ARGB words (reinterpreted). Device pointers, async on `stream`. */
int c2rt_deinterleave(const void* gathered, void* frame, uint32_t width, uint32_t height, uint32_t elem_words,
                      uint32_t n_ranks, uint32_t band_rows, uint32_t rows_per_rank_padded, void* stream);

/* Replaces renderPixel (renderer.d:46-57): one un-antialiased sample at the pixel corner. */
int c2rt_render_pixel(c2rt_scene* scene, const c2rt_camera* camera, const c2rt_settings* settings,
                      int x, int y, float rgb[3], c2rt_hit* hit);

/* Number of rows rank `rank` owns in a frame of `height` rows. */
uint32_t c2rt_band_rows_owned(uint32_t height, uint32_t rank, uint32_t n_ranks, uint32_t band_rows);

/* The pinned generator that stands in for util/random.d's libc rand(): 31 bits keyed
 * by (seed, pixel, AA tap, DOF sample, draw index); uniform(0,1) = value / 2147483647.0 . */
uint32_t c2rt_rng_u31(uint64_t seed, uint32_t px, uint32_t py, uint32_t tap, uint32_t sample, uint32_t draw);

/* The 4097-entry sRGB compression table (color.d:209-229) used for ARGB output. */
void c2rt_srgb_table(uint8_t out[4097]);

/* Peer-visible device buffers for the one-process-per-GPU path: rank 0 allocates the frame and
 * exports a 64-byte handle; the other ranks import it and pass the mapped pointer as d_rgb so their
 * kernels store bands directly into rank 0's memory over NVLink. */
int c2rt_frame_alloc(size_t bytes, void** d_ptr);   /* zero-filled */
int c2rt_frame_free(void* d_ptr);
int c2rt_frame_export(void* d_ptr, uint8_t handle[64]);
int c2rt_frame_import(const uint8_t handle[64], void** d_ptr);
int c2rt_frame_unimport(void* d_ptr);
/* asynchronous byte fill of (part of) such a frame on `stream` (e.g. 0xFF = NaN pixels before a checked frame, so that a band
 * that never arrived cannot pass for one left over from the previous frame) */
int c2rt_frame_memset(void* d_ptr, int byte_value, size_t bytes, void* stream);
/* asynchronous device->host copy of (part of) such a frame on `stream` (host memory should be pinned) */
int c2rt_frame_download(void* host_dst, const void* d_src, size_t bytes, void* stream);

/* Device-side start gate for the one-process-per-GPU path: every rank enqueues c2rt_gate(flags, n_ranks, round, stream) in
 * front of a frame, with round = 1, 2, 3, ... counted per flags allocation; the one-thread kernel adds 1 to flags[0] (rank 0's
 * memory, over NVLink for the peers) and waits until it reaches round * n_ranks, i.e. until every rank arrived.  The ranks'
 * render kernels then start together without a host round trip, and no peer can store bands of frame k + 1 while rank 0's
 * stream still works on (renders, copies out, clears) frame k.  Same ~2 s time-out rule and error word (flags[n_ranks]) as
 * c2rt_band.done_flags. */
int c2rt_gate(void* d_flags, uint32_t n_ranks, uint32_t round, void* stream);

/* Page-locks a caller-owned host buffer (e.g. the D host's Image!Color.pixels, which is ordinary GC memory) so
 * that c2rt_render's device->host copies run asynchronously at full PCIe rate and overlap with rendering.
 * Optional: without it the copies still work, through the driver's staging path.  Unpin before freeing. */
int c2rt_pin_host_buffer(void* ptr, size_t bytes);
int c2rt_unpin_host_buffer(void* ptr);

/* Micro-benchmarks used by bench.py to measure the roofline denominators on the box:
 * dependent-free FFMA / DFMA throughput in TFLOP/s on the current device. */
int c2rt_measure_fma_peak(int fp64, double* tflops, double* sm_clock_mhz_est);

/* Test hook (no GPU needed): drives the helper-thread pool c2rt_render uses after c2rt_init(N > 1) through
 * `rounds` frames of dummy work on `workers` threads, with the spin and the sleep hand-off paths both
 * exercised, restarting the pool in between.  Returns the number of work items executed
 * (== workers * rounds when nothing was lost or duplicated), negative on bad arguments. */
long long c2rt_selftest_device_pool(int workers, int rounds);

#ifdef __cplusplus
}
#endif
#endif /* C2RT_H */
