// Reference-side binding for libc2rt.so (see INTEGRATION.md).  Written against the reference checkout;
// NOT compiled in this repository's image (no D toolchain: SURVEY.md F2) — the C++ mirror in
// chess2rt_b200/host/ is the tested twin of this file.
module rt.cuda_backend;

extern (C) @nogc nothrow:

enum C2RT_ABI_VERSION = 2;
enum : int { C2RT_CANCELLED = 1, C2RT_OK = 0, C2RT_ERR_INVALID_ARG = -1, C2RT_ERR_UNSUPPORTED = -2, C2RT_ERR_CUDA = -3,
             C2RT_ERR_NOT_INITIALISED = -4, C2RT_ERR_LIMIT = -5 }
enum : int { C2RT_GEOM_PLANE, C2RT_GEOM_SPHERE, C2RT_GEOM_CUBE, C2RT_GEOM_CSG_UNION, C2RT_GEOM_CSG_INTER, C2RT_GEOM_CSG_DIFF }
enum : int { C2RT_SHADER_LAMBERT, C2RT_SHADER_PHONG }
enum : int { C2RT_TEX_CHECKER, C2RT_TEX_PROCEDURE2, C2RT_TEX_BITMAP }
enum : int { C2RT_ENV_BLACK, C2RT_ENV_CUBEMAP }      // CUBEMAP: extension, environment.d has the black stub only

struct c2rt_scene_desc
{
    uint struct_size, abi_version;
    uint n_nodes;    const(int)* node_geom, node_shader;
                     const(double)* node_transform, node_inverse, node_inverse_t, node_offset;
    uint n_geoms;    const(int)* geom_type; const(double)* geom_params; const(int)* geom_left, geom_right;
    uint n_shaders;  const(int)* shader_type; const(float)* shader_color; const(int)* shader_texture;
                     const(double)* shader_exponent; const(float)* shader_strength;
    uint n_textures; const(int)* tex_type; const(float)* tex_colors; const(double)* tex_params;
                     const(int)* tex_width, tex_height; const(ulong)* tex_texel_offset;
                     const(float)* texels; ulong n_texels;
    uint n_lights;   const(double)* light_pos; const(float)* light_color, light_power;
    int env_type, env_reserved;                       // C2RT_ENV_*
    int[6] env_face_width, env_face_height;           // +x, -x, +y, -y, +z, -z
    ulong[6] env_face_texel_offset;                   // into `texels`
}

struct c2rt_camera
{
    double[3] pos, up_left, up_right, down_left, right_dir, up_dir, front_dir;
    uint frame_width, frame_height;
    int dof; uint num_samples;
    double focal_plane_dist, disc_multiplier, stereo_separation;
}

struct c2rt_settings
{
    uint frame_width, frame_height;
    int aa_enabled, gi_enabled, prepass_enabled, prepass_only;
    uint max_trace_depth;
    float[3] ambient_light;
    ulong rng_seed;
    int count_rays; uint bucket_size;
    uint paths_per_pixel, reserved;
}

struct c2rt_stats { double kernel_ms, total_ms; ulong primary_rays, shadow_rays; uint n_gpus, launches; }
struct c2rt_hit   { int node, reserved; double dist; double[3] p, normal; double u, v; }
struct c2rt_scene;   // opaque

int  c2rt_init(int n_gpus, const(int)* device_ids);
void c2rt_shutdown();
const(char)* c2rt_last_error();
int  c2rt_scene_create(const(c2rt_scene_desc)* desc, c2rt_scene** out_);
void c2rt_scene_destroy(c2rt_scene* scene);
int  c2rt_render(c2rt_scene* scene, const(c2rt_camera)* cam, const(c2rt_settings)* set,
                 float* rgb, uint* argb, c2rt_stats* stats);
int  c2rt_cancel();                                  // any thread, while another one is inside c2rt_render
int  c2rt_render_pixel(c2rt_scene* scene, const(c2rt_camera)* cam, const(c2rt_settings)* set,
                       int x, int y, float* rgb, c2rt_hit* hit);
int  c2rt_pin_host_buffer(void* ptr, size_t bytes);
int  c2rt_unpin_host_buffer(void* ptr);
