// Scene flattener: lowers the host object graph (rt::Scene and the CSG trees hanging off its
// geometries) into the structure-of-arrays description of include/c2rt.h.  This is subsystem (1)
// of the north star; field map in SURVEY.md Appendix B.  Object references become indices;
// objects shared by several nodes keep a single entry; CSG children keep their identity as an index
// (the reference's `current.g is left` test, /root/reference/source/rt/geometry.d:314).
// Runs once per scene (load time); the camera / settings blocks are refreshed per frame.
#pragma once
#include <map>
#include <vector>

#include "../../include/c2rt.h"
#include "rt.hpp"

namespace rt {

struct FlatScene {
    std::vector<int32_t> node_geom, node_shader;
    std::vector<double> node_transform, node_inverse, node_inverse_t, node_offset;
    std::vector<int32_t> geom_type, geom_left, geom_right;
    std::vector<double> geom_params;
    std::vector<int32_t> shader_type, shader_texture;
    std::vector<float> shader_color, shader_strength;
    std::vector<double> shader_exponent;
    std::vector<int32_t> tex_type, tex_width, tex_height;
    std::vector<float> tex_colors;
    std::vector<double> tex_params;
    std::vector<uint64_t> tex_texel_offset;
    std::vector<float> texels;
    std::vector<double> light_pos;
    std::vector<float> light_color, light_power;
    int32_t env_type = C2RT_ENV_BLACK;
    int32_t env_face_width[6] = {0}, env_face_height[6] = {0};
    uint64_t env_face_texel_offset[6] = {0};

    c2rt_scene_desc desc() const;  // borrows the vectors above
};

FlatScene flatten(const Scene& scene);                        // throws InvalidSceneException on dangling references
c2rt_camera flattenCamera(const Camera& cam);                 // call after Camera.beginFrame
c2rt_settings flattenSettings(const GlobalSettings& s, uint64_t rngSeed = 0, bool countRays = false);

}  // namespace rt
