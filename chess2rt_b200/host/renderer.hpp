// The renderer seam, re-pointed at libc2rt.so.  Same entry points and argument meaning as
// /root/reference/source/rt/renderer.d:
//   renderSceneAsync :23-44   scene.beginFrame(); render on a spawned thread; clears *isRendering when done
//   renderPixel      :46-57   one un-antialiased sample + TraceResult for the debug pixel pick
//   Renderer         :59-189  Renderer(scene, output[, isRendering, isStopRequested]).renderRT()
// The body of renderRT (bucket list, prepass, 1-spp pass, AA pass over a CPU TaskPool) is replaced
// by one call of c2rt_render; the prepass and the dead needsAA mask have no pixel effect and are
// dropped (SURVEY.md §7 "dead work").  Errors from the backend surface as rt::BackendException;
// there is no CPU fallback.
#pragma once
#include <atomic>
#include <thread>
#include <tuple>

#include "flatten.hpp"
#include "rt.hpp"

namespace rt {

struct TraceResult {  // renderer.d:14-21 (closestNode as an index into scene.nodes, -1 = miss)
    int closestNode = -1;
    double dist = 1e99;
    Vector p, normal;
    double u = NAN, v = NAN;
    bool hitLight = false;
    Color hitLightColor;
};

struct RenderOptions {        // knobs the reference does not have
    uint64_t rngSeed = 0;     // seed of the pinned DOF generator
    bool countRays = false;   // fill lastStats.primary_rays / shadow_rays
    Image<uint32_t>* argb = nullptr;  // optional Color.toRGB32 plane produced on the GPU
    bool argbOnly = false;    // interactive hosts that only blit `argb` (sdl2_gui.d:139-155): the float frame is neither written nor copied
};

struct Renderer {
    Renderer(const Scene& scene, Image<Color>& output, std::atomic<bool>* isRendering = nullptr,
             const std::atomic<bool>* isStopRequested = nullptr);

    void renderRT();                               // renderer.d:83-189
    Color renderPixelNoAA(int x, int y);           // renderer.d:223-228 (also records lastTracingResult)

    TraceResult lastTracingResult;
    c2rt_stats lastStats{};
    RenderOptions options;
    bool cancelled = false;                        // the last renderRT ended early on a stop request (the frame is partial)

private:
    const Scene& scene_;
    Image<Color>& output_;
    std::atomic<bool>* isRendering_;
    const std::atomic<bool>* isStopRequested_;
    c2rt_scene* device();
};

// Spawns the render thread like the reference does; `worker` (optional) receives the thread so a
// headless caller can join it instead of polling *isRendering.
void renderSceneAsync(Scene& scene, Image<Color>& output, std::atomic<bool>* isRendering,
                      const std::atomic<bool>* needsRendering, std::thread* worker = nullptr,
                      const RenderOptions& options = RenderOptions());

std::tuple<Color, TraceResult> renderPixel(Scene& scene, Image<Color>& output, int x, int y);

// The stop request of the reference (raytracer_demo.d:102-124 sets needsRendering; renderer.d:93-97,129,147,180 polls it between
// passes) for a frame that is already on the GPU: sets *isStopRequested like the GUI does and tells the backend, which skips
// every tile that has not started.  Callable from the GUI thread while the render thread is inside renderRT.
void requestStop(std::atomic<bool>* isStopRequested);

// c2rt_init wrapper: selects the GPUs the next frames are banded over (default: device 0 only).
void setRenderDevices(int nGpus, const int* deviceIds = nullptr);

}  // namespace rt
