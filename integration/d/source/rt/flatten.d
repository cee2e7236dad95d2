// Reference-side binding for libc2rt.so (see INTEGRATION.md).  Written against the reference checkout;
// NOT compiled in this repository's image (no D toolchain: SURVEY.md F2) — the C++ mirror in
// chess2rt_b200/host/ is the tested twin of this file.
module rt.flatten;

import rt.scene, rt.node, rt.geometry, rt.shader, rt.texture, rt.light, rt.camera, rt.globalsettings;
import rt.cuda_backend;
import std.algorithm : countUntil;

struct FlatScene
{
    int[] nodeGeom, nodeShader; double[] nodeM, nodeMinv, nodeMinvT, nodeOffset;
    int[] geomType, geomLeft, geomRight; double[] geomParams;
    int[] shaderType, shaderTex; float[] shaderColor, shaderStrength; double[] shaderExponent;
    int[] texType, texW, texH; float[] texColors; double[] texParams; ulong[] texOffset; float[] texels;
    double[] lightPos; float[] lightColor, lightPower;
    int envType = C2RT_ENV_BLACK; int[6] envW, envH; ulong[6] envOffset;

    c2rt_scene_desc desc() const
    {
        c2rt_scene_desc d;
        d.struct_size = c2rt_scene_desc.sizeof;  d.abi_version = C2RT_ABI_VERSION;
        d.n_nodes = cast(uint) nodeGeom.length;
        d.node_geom = nodeGeom.ptr;  d.node_shader = nodeShader.ptr;
        d.node_transform = nodeM.ptr; d.node_inverse = nodeMinv.ptr; d.node_inverse_t = nodeMinvT.ptr; d.node_offset = nodeOffset.ptr;
        d.n_geoms = cast(uint) geomType.length;
        d.geom_type = geomType.ptr; d.geom_params = geomParams.ptr; d.geom_left = geomLeft.ptr; d.geom_right = geomRight.ptr;
        d.n_shaders = cast(uint) shaderType.length;
        d.shader_type = shaderType.ptr; d.shader_color = shaderColor.ptr; d.shader_texture = shaderTex.ptr;
        d.shader_exponent = shaderExponent.ptr; d.shader_strength = shaderStrength.ptr;
        d.n_textures = cast(uint) texType.length;
        d.tex_type = texType.ptr; d.tex_colors = texColors.ptr; d.tex_params = texParams.ptr;
        d.tex_width = texW.ptr; d.tex_height = texH.ptr; d.tex_texel_offset = texOffset.ptr;
        d.texels = texels.ptr; d.n_texels = texels.length / 3;
        d.n_lights = cast(uint) lightPower.length;
        d.light_pos = lightPos.ptr; d.light_color = lightColor.ptr; d.light_power = lightPower.ptr;
        d.env_type = envType; d.env_face_width = envW; d.env_face_height = envH; d.env_face_texel_offset = envOffset;
        return d;
    }
}

FlatScene flatten(const Scene s)
{
    FlatScene f;
    int geomIndex(const Geometry g) { return cast(int) s.geometries.countUntil!(x => x is g); }

    foreach (g; s.geometries)
    {
        double[4] p = 0; int type, l = -1, r = -1;
        if (auto pl = cast(const Plane) g)       { type = C2RT_GEOM_PLANE;  p[0] = pl.y; p[1] = pl.limit; }
        else if (auto sp = cast(const Sphere) g) { type = C2RT_GEOM_SPHERE; p[0 .. 3] = sp.getCenter.v[]; p[3] = sp.getR; }
        else if (auto cu = cast(const Cube) g)   { type = C2RT_GEOM_CUBE;   p[0 .. 3] = cu.center.v[];   p[3] = cu.side; }
        else if (auto op = cast(const CsgOp) g)
        {
            type = cast(const CsgUnion) g ? C2RT_GEOM_CSG_UNION : cast(const CsgInter) g ? C2RT_GEOM_CSG_INTER : C2RT_GEOM_CSG_DIFF;
            l = geomIndex(op.getLeft); r = geomIndex(op.getRight);
        }
        f.geomType ~= type; f.geomLeft ~= l; f.geomRight ~= r; f.geomParams ~= p[];
    }
    foreach (t; s.textures)
    {
        float[18] c = 0; double[6] p = 0; int type, w, h; ulong off;
        if (auto ch = cast(const Checker) t)
        {   type = C2RT_TEX_CHECKER; c[0 .. 3] = ch.color1.components[]; c[3 .. 6] = ch.color2.components[]; p[0] = ch.size; }
        else if (auto pr = cast(const Procedure2) t)
        {
            type = C2RT_TEX_PROCEDURE2;
            foreach (k; 0 .. 3) { c[3*k .. 3*k+3] = pr.colorU[k].components[]; c[9+3*k .. 12+3*k] = pr.colorV[k].components[];
                                  p[k] = pr.freqU[k]; p[3+k] = pr.freqV[k]; }
        }
        else if (auto bm = cast(const BitmapTexture) t)
        {
            type = C2RT_TEX_BITMAP; w = cast(int) bm.getBitmap.width; h = cast(int) bm.getBitmap.height;
            p[0] = bm.getScaling;                       // float widened exactly, as `u *= scaling` does (texture.d:118)
            off = f.texels.length / 3;
            foreach (px; bm.getBitmap.data.pixels) f.texels ~= px.components[];   // post-gamma Image!Color (texture.d:137-141)
        }
        f.texType ~= type; f.texW ~= w; f.texH ~= h; f.texOffset ~= off; f.texColors ~= c[]; f.texParams ~= p[];
    }
    foreach (sh; s.shaders)
    {
        int type, tex = -1; double e = 0; float st = 0;
        if (auto la = cast(const Lambert) sh) { type = C2RT_SHADER_LAMBERT; if (la.texture) tex = cast(int) s.textures.countUntil!(x => x is la.texture); }
        else if (auto ph = cast(const Phong) sh) { type = C2RT_SHADER_PHONG; e = ph.exponent; st = ph.strength;
                                                    if (ph.texture) tex = cast(int) s.textures.countUntil!(x => x is ph.texture); }
        f.shaderType ~= type; f.shaderTex ~= tex; f.shaderColor ~= sh.color.components[]; f.shaderExponent ~= e; f.shaderStrength ~= st;
    }
    foreach (l; s.lights)
    {
        auto pl = cast(const PointLight) l;            // the only Light subclass (light.d:52)
        f.lightPos ~= pl.pos.v[]; f.lightColor ~= pl.lightColor.components[]; f.lightPower ~= pl.lightPower;
    }
    // cubemap-environment extension (INTEGRATION.md section 8): Environment gains `Bitmap[6] faces; bool cubemap`
    if (s.environment.cubemap)
    {
        f.envType = C2RT_ENV_CUBEMAP;
        foreach (k, ref face; s.environment.faces)
        {
            f.envW[k] = cast(int) face.width; f.envH[k] = cast(int) face.height; f.envOffset[k] = f.texels.length / 3;
            foreach (px; face.data.pixels) f.texels ~= px.components[];
        }
    }
    foreach (n; s.nodes)
    {
        f.nodeGeom ~= geomIndex(n.geom);
        f.nodeShader ~= cast(int) s.shaders.countUntil!(x => x is n.shader);
        auto t = n.transform;                          // Transform{transform, inverseTransform, transposedInverse, offset}
        foreach (r; 0 .. 3) foreach (c; 0 .. 3) { f.nodeM ~= t.getTransform.c[r][c]; f.nodeMinv ~= t.getInverse.c[r][c]; f.nodeMinvT ~= t.getTransposedInverse.c[r][c]; }
        f.nodeOffset ~= t.getOffset.v[];
    }
    return f;
}

c2rt_camera flattenCamera(const Camera c)              // call after Camera.beginFrame (camera.d:77-117)
{
    c2rt_camera o;
    o.pos = c.pos.v; o.up_left = c.getUpLeft.v; o.up_right = c.getUpRight.v; o.down_left = c.getDownLeft.v;
    o.right_dir = c.getRightDir.v; o.up_dir = c.getUpDir.v; o.front_dir = c.getFrontDir.v;
    o.frame_width = cast(uint) c.frameWidth; o.frame_height = cast(uint) c.frameHeight;
    o.dof = c.dof; o.num_samples = cast(uint) c.numSamples;
    o.focal_plane_dist = c.focalPlaneDist; o.disc_multiplier = c.discMultiplier; o.stereo_separation = c.stereoSeparation;
    return o;
}

c2rt_settings flattenSettings(const GlobalSettings s, ulong seed = 0)
{
    c2rt_settings o;
    o.frame_width = s.frameWidth; o.frame_height = s.frameHeight;
    o.aa_enabled = s.AAEnabled; o.gi_enabled = s.GIEnabled; o.prepass_enabled = s.prepassEnabled; o.prepass_only = s.prepassOnly;
    o.max_trace_depth = s.maxTraceDepth; o.ambient_light = s.ambientLightColor.components; o.rng_seed = seed;
    o.bucket_size = s.bucketSize; o.paths_per_pixel = s.pathsPerPixel;
    return o;
}
