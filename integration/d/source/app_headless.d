// Reference-side addition for source/app.d (INTEGRATION.md section 6): the headless mode the north star asks for.  Written against
// the reference checkout; NOT compiled in this repository's image (no D toolchain: SURVEY.md F2) — the tested twin is
// chess2rt_b200/host/main.cpp (same flags, same flow).
//
// The edit of source/app.d itself is three hunks (the rest of the file stays as it is):
//
//   @@ main(): after `string sceneFilePath = "";`                                  (app.d:13)
//   +    bool headless = false, noDof = false, noAA = false;
//   +    string outPath = "";
//   +    uint width = 0, height = 0, gpus = 1, orbit = 0;
//   +    ulong seed = 0;
//   @@ main(): the getopt call                                                       (app.d:15)
//   -    getopt(args, "file", &sceneFilePath);
//   +    getopt(args, "file", &sceneFilePath,
//   +                 "headless", &headless, "out", &outPath, "width", &width, "height", &height,
//   +                 "gpus", &gpus, "no-dof", &noDof, "no-aa", &noAA, "seed", &seed, "orbit", &orbit);
//   +    if (headless)
//   +        return runHeadless(sceneFilePath, outPath, width, height, gpus, noDof, noAA, seed, orbit);
//   @@ imports
//   +import app_headless : runHeadless;
module app_headless;

import std.stdio : writefln;

/// The flow of RTDemo.resetScene + render + takeScreenshot (raytracer_demo.d:145-187, 102-124, 227-238) without a window.
void runHeadless(string scenePath, string outPath, uint width, uint height, uint gpus, bool noDof, bool noAA, ulong seed, uint orbit)
{
    import std.datetime.stopwatch : StopWatch, AutoStart;
    import imageio.image : Image;
    import rt.sceneloader : parseSceneFromFile;
    import rt.bitmap : Bitmap;
    import rt.color : Color;
    import rt.cuda_backend, rt.flatten, rt.renderer_cuda;

    c2rt_init(cast(int) gpus, null);                             // frames are banded over `gpus` devices from here on
    scope (exit) c2rt_shutdown();

    auto scene = parseSceneFromFile(scenePath);                  // scene_loader.d:20, unchanged
    if (width && height)                                         // "as if the scene file had been edited" (camera.d:231-236,254)
    {
        scene.settings.frameWidth = width; scene.settings.frameHeight = height;
        scene.camera.setFrameSize(width, height);
    }
    if (noDof) scene.camera.dof = false;
    if (noAA) scene.settings.AAEnabled = false;

    Image!Color screen;
    screen.alloc(scene.settings.frameWidth, scene.settings.frameHeight);       // raytracer_demo.d:181-182
    c2rt_pin_host_buffer(screen.pixels.ptr, screen.pixels.length * Color.sizeof);
    scope (exit) c2rt_unpin_host_buffer(screen.pixels.ptr);

    auto sw = StopWatch(AutoStart.yes);
    renderSceneSync(scene, screen, seed);                        // renderSceneAsync's body without the spawn (renderer_cuda.d)
    writefln("%sx%s on %s GPU(s): %s ms end to end", screen.w, screen.h, gpus, sw.peek.total!"usecs" / 1000.0);

    if (outPath.length)
        (const Bitmap(screen)).saveImage(outPath);               // bitmap.d:84-103 -> bmp.d:195-237 (rows unpadded)

    if (orbit)                                                   // the interactive loop without a window: camera-only updates
    {
        sw.reset();
        foreach (k; 0 .. orbit)
        {
            scene.camera.rotate(360.0 / orbit, 0, 0);            // camera.d:211-229, what the arrow keys do
            renderSceneSync(scene, screen, seed);
        }
        writefln("orbit: %s frames, %s ms per frame", orbit, sw.peek.total!"usecs" / 1000.0 / orbit);
    }
}
