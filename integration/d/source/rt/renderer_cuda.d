// Reference-side binding for libc2rt.so (see INTEGRATION.md).  Written against the reference checkout;
// NOT compiled in this repository's image (no D toolchain: SURVEY.md F2) — the C++ mirror in
// chess2rt_b200/host/ is the tested twin of this file.
// Replacement bodies for renderSceneAsync / renderPixel of source/rt/renderer.d:23-57.
module rt.renderer_cuda;

import std.concurrency : spawn;
import std.exception : enforce;
import std.stdio : stderr;
import std.string : fromStringz;
import std.typecons : tuple;
import core.atomic : atomicLoad, atomicStore;
import imageio.image : Image;
import rt.scene, rt.color, rt.importedtypes, rt.exception, rt.renderer : TraceResult;
import rt.cuda_backend, rt.flatten;

// new: one uploaded scene per Scene object, created on first use
private c2rt_scene* deviceScene(Scene scene)
{
    if (scene.gpuHandle is null)                      // new field `c2rt_scene* gpuHandle` in class Scene
    {
        auto flat = flatten(scene);                   // flat must outlive the call only: the library deep-copies
        auto d = flat.desc();
        enforce!RTException(c2rt_scene_create(&d, &scene.gpuHandle) == C2RT_OK, c2rt_last_error().fromStringz.idup);
    }
    return scene.gpuHandle;
}

void renderSceneAsync(Scene scene, Image!Color output, shared(bool)* isRendering, const shared(bool)* needsRendering)
{
    scene.beginFrame();                                // unchanged (renderer.d:31)
    spawn((shared Scene s, shared Image!Color o, shared(bool)* isWorking, const shared(bool)* isStopping)
    {
        scope (exit) if (isWorking !is null) (*isWorking).atomicStore(false);     // renderer.d:87-91
        if (isStopping !is null && (*isStopping).atomicLoad()) return;            // renderer.d:93-97
        auto sc = cast() s; auto img = cast() o;
        auto cam = flattenCamera(sc.camera);
        auto set = flattenSettings(sc.settings);
        // Image!Color.pixels is a tightly packed float[3] array, row-major, top row first (imageio/image.d:18-54)
        auto rc = c2rt_render(deviceScene(sc), &cam, &set, cast(float*) img.pixels.ptr, null, null);
        // C2RT_CANCELLED: a stop request (requestStop below) reached the frame in flight — the reference's `return end()`
        if (rc != C2RT_OK && rc != C2RT_CANCELLED) stderr.writeln("c2rt_render: ", c2rt_last_error().fromStringz);
    }, cast(shared) scene, cast(shared) output, isRendering, needsRendering);
}

// new: the synchronous form app.d's headless mode uses (same body as the spawned delegate above)
void renderSceneSync(Scene scene, Image!Color output, ulong seed = 0)
{
    scene.beginFrame();
    auto cam = flattenCamera(scene.camera);
    auto set = flattenSettings(scene.settings, seed);
    enforce!RTException(c2rt_render(deviceScene(scene), &cam, &set, cast(float*) output.pixels.ptr, null, null) == C2RT_OK,
                        c2rt_last_error().fromStringz.idup);
}

// new: what RTDemo calls where it sets `needsRendering = true` while a frame is being rendered (raytracer_demo.d:85-124,268-340),
// so that the stop request the reference polls between passes (renderer.d:93-97,129,147,180) also reaches a frame already on
// the GPU: tiles that have not started are skipped and c2rt_render returns C2RT_CANCELLED.
void requestStop(shared(bool)* needsRendering)
{
    if (needsRendering !is null) (*needsRendering).atomicStore(true);
    c2rt_cancel();
}

auto renderPixel(Scene scene, Image!Color output, int x, int y)
{
    scene.beginFrame();                                // renderer.d:50
    auto cam = flattenCamera(scene.camera); auto set = flattenSettings(scene.settings);
    float[3] rgb; c2rt_hit hit;
    enforce!RTException(c2rt_render_pixel(deviceScene(scene), &cam, &set, x, y, rgb.ptr, &hit) == C2RT_OK,
                        c2rt_last_error().fromStringz.idup);
    auto color = Color(rgb[0], rgb[1], rgb[2]);
    output[x, y] = color;                              // renderer.d:226
    TraceResult result;                                // renderer.d:14-21
    result.closestNode = hit.node >= 0 ? scene.nodes[hit.node] : null;
    result.data.dist = hit.dist; result.data.p = Vector(hit.p); result.data.normal = Vector(hit.normal);
    result.data.u = hit.u; result.data.v = hit.v;
    return tuple(color, result);
}
