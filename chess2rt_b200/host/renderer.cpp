#include "renderer.hpp"

#include <cstring>

namespace rt {
namespace {

[[noreturn]] void raise(int status) { throw BackendException(status, c2rt_last_error()); }

struct DeviceSceneHolder {
    c2rt_scene* handle = nullptr;
    ~DeviceSceneHolder() { c2rt_scene_destroy(handle); }
};

}  // namespace

void setRenderDevices(int nGpus, const int* deviceIds) {
    int rc = c2rt_init(nGpus, deviceIds);
    if (rc != C2RT_OK) raise(rc);
}

Renderer::Renderer(const Scene& scene, Image<Color>& output, std::atomic<bool>* isRendering, const std::atomic<bool>* isStopRequested)
    : scene_(scene), output_(output), isRendering_(isRendering), isStopRequested_(isStopRequested) {}

c2rt_scene* Renderer::device() {
    if (!scene_.deviceScene) {
        FlatScene flat = flatten(scene_);
        c2rt_scene_desc d = flat.desc();
        auto holder = std::make_shared<DeviceSceneHolder>();
        int rc = c2rt_scene_create(&d, &holder->handle);
        if (rc != C2RT_OK) raise(rc);
        scene_.deviceScene = holder;
    }
    return static_cast<DeviceSceneHolder*>(scene_.deviceScene.get())->handle;
}

void Renderer::renderRT() {
    struct Done {  // renderer.d:87-91 end()
        std::atomic<bool>* flag;
        ~Done() { if (flag) flag->store(false); }
    } done{isRendering_};

    const uint32_t W = scene_.settings.frameWidth, H = scene_.settings.frameHeight;
    if (output_.width != W || output_.height != H || output_.pixels.size() < (size_t)W * H)
        throw RTException("output image size does not match settings.frameWidth x frameHeight");
    if (isStopRequested_ && isStopRequested_->load()) return;  // renderer.d:93-97,129

    c2rt_camera cam = flattenCamera(scene_.camera);
    c2rt_settings set = flattenSettings(scene_.settings, options.rngSeed, options.countRays);
    uint32_t* argb = nullptr;
    if (options.argb) {
        options.argb->alloc(W, H);
        argb = options.argb->pixels.data();
    }
    static_assert(sizeof(Color) == 3 * sizeof(float), "Color must be three packed floats (color.d:27-35)");
    if (options.argbOnly && !argb) throw RTException("argbOnly needs RenderOptions::argb");
    float* rgb = options.argbOnly ? nullptr : reinterpret_cast<float*>(output_.pixels.data());
    int rc = c2rt_render(device(), &cam, &set, rgb, argb, &lastStats);
    cancelled = rc == C2RT_CANCELLED;   // a stop request reached the frame in flight: return like the reference's `return end()`
    if (rc != C2RT_OK && !cancelled) raise(rc);
}

void requestStop(std::atomic<bool>* isStopRequested) {
    if (isStopRequested) isStopRequested->store(true);
    c2rt_cancel();
}

Color Renderer::renderPixelNoAA(int x, int y) {
    c2rt_camera cam = flattenCamera(scene_.camera);
    c2rt_settings set = flattenSettings(scene_.settings, options.rngSeed, false);
    float rgb[3];
    c2rt_hit hit;
    int rc = c2rt_render_pixel(device(), &cam, &set, x, y, rgb, &hit);
    if (rc != C2RT_OK) raise(rc);
    Color c(rgb[0], rgb[1], rgb[2]);
    output_(x, y) = c;  // renderer.d:226
    lastTracingResult = TraceResult();
    lastTracingResult.closestNode = hit.node;
    lastTracingResult.dist = hit.dist;
    if (hit.node >= 0) {
        lastTracingResult.p = Vector(hit.p[0], hit.p[1], hit.p[2]);
        lastTracingResult.normal = Vector(hit.normal[0], hit.normal[1], hit.normal[2]);
        lastTracingResult.u = hit.u;
        lastTracingResult.v = hit.v;
    }
    return c;
}

void renderSceneAsync(Scene& scene, Image<Color>& output, std::atomic<bool>* isRendering, const std::atomic<bool>* needsRendering,
                      std::thread* worker, const RenderOptions& options) {
    scene.beginFrame();  // renderer.d:31
    std::thread t([&scene, &output, isRendering, needsRendering, options]() {
        Renderer renderer(scene, output, isRendering, needsRendering);
        renderer.options = options;
        try {
            renderer.renderRT();
        } catch (const std::exception& e) {
            // the reference's render thread has nowhere to report to either; keep the message visible
            fprintf(stderr, "render thread: %s\n", e.what());
        }
    });
    if (worker) *worker = std::move(t);
    else t.detach();
}

std::tuple<Color, TraceResult> renderPixel(Scene& scene, Image<Color>& output, int x, int y) {
    scene.beginFrame();  // renderer.d:50
    Renderer renderer(scene, output);
    Color color = renderer.renderPixelNoAA(x, y);
    return std::make_tuple(color, renderer.lastTracingResult);
}

}  // namespace rt
