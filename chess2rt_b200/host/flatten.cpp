#include "flatten.hpp"

#include <cstring>

namespace rt {
namespace {

template <class T>
int indexOf(const std::vector<std::unique_ptr<T>>& v, const T* p) {
    for (size_t i = 0; i < v.size(); i++)
        if (v[i].get() == p) return (int)i;
    return -1;
}

void pushMatrix(std::vector<double>& dst, const Matrix& m) {
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) dst.push_back(m.c[r][c]);
}

}  // namespace

FlatScene flatten(const Scene& scene) {
    FlatScene f;

    // geometries, in declaration order (CSG children are always earlier: geometry.d:343-347)
    for (size_t i = 0; i < scene.geometries.size(); i++) {
        const Geometry* g = scene.geometries[i].get();
        double p[4] = {0, 0, 0, 0};
        int type, left = -1, right = -1;
        if (auto* pl = dynamic_cast<const Plane*>(g)) {
            type = C2RT_GEOM_PLANE;
            p[0] = pl->y;
            p[1] = pl->limit;
        } else if (auto* sp = dynamic_cast<const Sphere*>(g)) {
            type = C2RT_GEOM_SPHERE;
            p[0] = sp->center.x; p[1] = sp->center.y; p[2] = sp->center.z; p[3] = sp->R;
        } else if (auto* cu = dynamic_cast<const Cube*>(g)) {
            type = C2RT_GEOM_CUBE;
            p[0] = cu->center.x; p[1] = cu->center.y; p[2] = cu->center.z; p[3] = cu->side;
        } else if (auto* op = dynamic_cast<const CsgOp*>(g)) {
            type = dynamic_cast<const CsgUnion*>(g) ? C2RT_GEOM_CSG_UNION : dynamic_cast<const CsgInter*>(g) ? C2RT_GEOM_CSG_INTER : C2RT_GEOM_CSG_DIFF;
            left = indexOf(scene.geometries, op->left);
            right = indexOf(scene.geometries, op->right);
            if (left < 0 || right < 0) throw InvalidSceneException("CSG geometry refers to a geometry outside the scene");
        } else {
            throw InvalidSceneException("unknown Geometry subclass");
        }
        f.geom_type.push_back(type);
        f.geom_left.push_back(left);
        f.geom_right.push_back(right);
        f.geom_params.insert(f.geom_params.end(), p, p + 4);
    }

    // textures
    for (size_t i = 0; i < scene.textures.size(); i++) {
        const Texture* t = scene.textures[i].get();
        float colors[18] = {0};
        double params[6] = {0};
        int type, w = 0, h = 0;
        uint64_t off = 0;
        if (auto* ch = dynamic_cast<const Checker*>(t)) {
            type = C2RT_TEX_CHECKER;
            colors[0] = ch->color1.r; colors[1] = ch->color1.g; colors[2] = ch->color1.b;
            colors[3] = ch->color2.r; colors[4] = ch->color2.g; colors[5] = ch->color2.b;
            params[0] = ch->size;
        } else if (auto* pr = dynamic_cast<const Procedure2*>(t)) {
            type = C2RT_TEX_PROCEDURE2;
            // texture.d:81 reads exactly three entries of each array (a D RangeError otherwise)
            if (pr->colorU.size() < 3 || pr->colorV.size() < 3 || pr->freqU.size() < 3 || pr->freqV.size() < 3)
                throw InvalidSceneException("Procedure2 needs three colorU / colorV / freqU / freqV entries");
            for (int k = 0; k < 3; k++) {
                colors[3 * k + 0] = pr->colorU[k].r; colors[3 * k + 1] = pr->colorU[k].g; colors[3 * k + 2] = pr->colorU[k].b;
                colors[9 + 3 * k + 0] = pr->colorV[k].r; colors[9 + 3 * k + 1] = pr->colorV[k].g; colors[9 + 3 * k + 2] = pr->colorV[k].b;
                params[k] = pr->freqU[k];
                params[3 + k] = pr->freqV[k];
            }
        } else if (auto* bm = dynamic_cast<const BitmapTexture*>(t)) {
            type = C2RT_TEX_BITMAP;
            w = (int)bm->bmp.width();
            h = (int)bm->bmp.height();
            params[0] = (double)bm->scaling;  // `u *= scaling` widens the float (texture.d:118)
            off = f.texels.size() / 3;
            for (const Color& c : bm->bmp.data.pixels) {
                f.texels.push_back(c.r);
                f.texels.push_back(c.g);
                f.texels.push_back(c.b);
            }
        } else {
            throw InvalidSceneException("unknown Texture subclass");
        }
        f.tex_type.push_back(type);
        f.tex_width.push_back(w);
        f.tex_height.push_back(h);
        f.tex_texel_offset.push_back(off);
        f.tex_colors.insert(f.tex_colors.end(), colors, colors + 18);
        f.tex_params.insert(f.tex_params.end(), params, params + 6);
    }

    // shaders
    for (size_t i = 0; i < scene.shaders.size(); i++) {
        const Shader* s = scene.shaders[i].get();
        int type, tex = -1;
        double exponent = 0;
        float strength = 0;
        if (auto* la = dynamic_cast<const Lambert*>(s)) {
            type = C2RT_SHADER_LAMBERT;
            if (la->texture) tex = indexOf(scene.textures, la->texture);
        } else if (auto* ph = dynamic_cast<const Phong*>(s)) {
            type = C2RT_SHADER_PHONG;
            if (ph->texture) tex = indexOf(scene.textures, ph->texture);
            exponent = ph->exponent;
            strength = ph->strength;
        } else {
            throw InvalidSceneException("unknown Shader subclass");
        }
        f.shader_type.push_back(type);
        f.shader_texture.push_back(tex);
        f.shader_color.push_back(s->color.r);
        f.shader_color.push_back(s->color.g);
        f.shader_color.push_back(s->color.b);
        f.shader_exponent.push_back(exponent);
        f.shader_strength.push_back(strength);
    }

    // lights (PointLight is the only Light subclass the reference has: light.d:52)
    for (auto& l : scene.lights) {
        auto* pl = dynamic_cast<const PointLight*>(l.get());
        if (!pl) throw InvalidSceneException("unknown Light subclass");
        f.light_pos.push_back(pl->pos.x); f.light_pos.push_back(pl->pos.y); f.light_pos.push_back(pl->pos.z);
        f.light_color.push_back(pl->lightColor.r); f.light_color.push_back(pl->lightColor.g); f.light_color.push_back(pl->lightColor.b);
        f.light_power.push_back(pl->lightPower);
    }

    // environment (cubemap extension: rt.hpp Environment): the six faces follow the bitmap textures' texels
    if (scene.environment.cubemap) {
        f.env_type = C2RT_ENV_CUBEMAP;
        for (int k = 0; k < 6; k++) {
            const Bitmap& b = scene.environment.faces[k];
            f.env_face_width[k] = (int32_t)b.width();
            f.env_face_height[k] = (int32_t)b.height();
            f.env_face_texel_offset[k] = f.texels.size() / 3;
            for (const Color& c : b.data.pixels) {
                f.texels.push_back(c.r);
                f.texels.push_back(c.g);
                f.texels.push_back(c.b);
            }
        }
    }

    // nodes
    for (auto& n : scene.nodes) {
        int g = indexOf(scene.geometries, n->geom), s = indexOf(scene.shaders, n->shader);
        if (g < 0 || s < 0) throw InvalidSceneException("node refers to a geometry / shader outside the scene");
        f.node_geom.push_back(g);
        f.node_shader.push_back(s);
        pushMatrix(f.node_transform, n->transform.transform);
        pushMatrix(f.node_inverse, n->transform.inverseTransform);
        pushMatrix(f.node_inverse_t, n->transform.transposedInverse);
        f.node_offset.push_back(n->transform.offset.x);
        f.node_offset.push_back(n->transform.offset.y);
        f.node_offset.push_back(n->transform.offset.z);
    }
    return f;
}

c2rt_scene_desc FlatScene::desc() const {
    c2rt_scene_desc d;
    memset(&d, 0, sizeof d);
    d.struct_size = sizeof d;
    d.abi_version = C2RT_ABI_VERSION;
    d.n_nodes = (uint32_t)node_geom.size();
    d.node_geom = node_geom.data();
    d.node_shader = node_shader.data();
    d.node_transform = node_transform.data();
    d.node_inverse = node_inverse.data();
    d.node_inverse_t = node_inverse_t.data();
    d.node_offset = node_offset.data();
    d.n_geoms = (uint32_t)geom_type.size();
    d.geom_type = geom_type.data();
    d.geom_params = geom_params.data();
    d.geom_left = geom_left.data();
    d.geom_right = geom_right.data();
    d.n_shaders = (uint32_t)shader_type.size();
    d.shader_type = shader_type.data();
    d.shader_color = shader_color.data();
    d.shader_texture = shader_texture.data();
    d.shader_exponent = shader_exponent.data();
    d.shader_strength = shader_strength.data();
    d.n_textures = (uint32_t)tex_type.size();
    d.tex_type = tex_type.data();
    d.tex_colors = tex_colors.data();
    d.tex_params = tex_params.data();
    d.tex_width = tex_width.data();
    d.tex_height = tex_height.data();
    d.tex_texel_offset = tex_texel_offset.data();
    d.texels = texels.data();
    d.n_texels = texels.size() / 3;
    d.n_lights = (uint32_t)light_power.size();
    d.light_pos = light_pos.data();
    d.light_color = light_color.data();
    d.light_power = light_power.data();
    d.env_type = env_type;
    for (int k = 0; k < 6; k++) {
        d.env_face_width[k] = env_face_width[k];
        d.env_face_height[k] = env_face_height[k];
        d.env_face_texel_offset[k] = env_face_texel_offset[k];
    }
    return d;
}

c2rt_camera flattenCamera(const Camera& cam) {
    c2rt_camera c;
    memset(&c, 0, sizeof c);
    auto put = [](double dst[3], const Vector& v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; };
    put(c.pos, cam.pos);
    put(c.up_left, cam.upLeft());
    put(c.up_right, cam.upRight());
    put(c.down_left, cam.downLeft());
    put(c.right_dir, cam.rightDir());
    put(c.up_dir, cam.upDir());
    put(c.front_dir, cam.frontDir());
    c.frame_width = (uint32_t)cam.frameWidth;
    c.frame_height = (uint32_t)cam.frameHeight;
    c.dof = cam.dof;
    c.num_samples = (uint32_t)cam.numSamples;
    c.focal_plane_dist = cam.focalPlaneDist;
    c.disc_multiplier = cam.discMultiplier;
    c.stereo_separation = cam.stereoSeparation;
    return c;
}

c2rt_settings flattenSettings(const GlobalSettings& s, uint64_t rngSeed, bool countRays) {
    c2rt_settings o;
    memset(&o, 0, sizeof o);
    o.frame_width = s.frameWidth;
    o.frame_height = s.frameHeight;
    o.aa_enabled = s.AAEnabled;
    o.gi_enabled = s.GIEnabled;
    o.prepass_enabled = s.prepassEnabled;
    o.prepass_only = s.prepassOnly;
    o.max_trace_depth = s.maxTraceDepth;
    o.ambient_light[0] = s.ambientLightColor.r;
    o.ambient_light[1] = s.ambientLightColor.g;
    o.ambient_light[2] = s.ambientLightColor.b;
    o.rng_seed = rngSeed;
    o.count_rays = countRays;
    o.bucket_size = s.bucketSize;
    o.paths_per_pixel = s.pathsPerPixel;
    return o;
}

}  // namespace rt
