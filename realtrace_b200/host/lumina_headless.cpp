// lumina_headless.cpp — what Serial/lumina.cpp does, minus the window: same includes, same scene
// set-up calls (lumina.cpp:302-370), same command line `[width] [height]` with even rounding
// (:480-486), one frame through RenderEngine, saved as PNG (own zlib writer, SaveImage) or binary PPM instead of going through DevIL (:424-439).
// It exists to show that host code written against the reference's headers builds and runs
// unchanged on top of the GPU core.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "camera.h"
#include "renderengine.h"
#include "world.h"
#include "material.h"
#include "object.h"
#include "sphere.h"
#include "triangle.h"
#include "cylinder.h"
#include "plane.h"
#include "lightsource.h"
#include "pointlightsource.h"

int main(int argc, char* argv[]) {
    int screen_width = 800, screen_height = 600;
    if (argc > 2) {
        screen_width = atoi(argv[1]);
        screen_height = atoi(argv[2]);
        screen_width -= (screen_width % 2);
        screen_height -= (screen_height % 2);
    }
    std::string obj = argc > 3 ? argv[3] : "assets/bob_tri.obj";
    std::string out = argc > 4 ? argv[4] : "lumina.ppm";
    try {
        Vector3D camera_position(60, 60, 0);
        Vector3D camera_target(0, 0, 0);
        Vector3D camera_up(0, 1, 0);
        float camera_fovy = 45;
        Camera* camera = new Camera(camera_position, camera_target, camera_up, camera_fovy, screen_width, screen_height);
        World* world = new World;
        world->setAmbient(Color(1));
        world->setBackground(Color(0.1, 0.3, 0.6));
        LightSource* light = new PointLightSource(world, Vector3D(0, 30, 30), Color(0.5, 1, 1));
        world->addLight(light);
        load_image_from_obj(world, obj);
        RenderEngine* engine = new RenderEngine(world, camera);
        while (!engine->renderLoop()) {}
        fprintf(stderr, "Rendering complete.\n");
        unsigned long long p, s, c;
        float ms;
        engine->frameStats(p, s, c, ms);
        fprintf(stderr, "%llu primary + %llu shadow + %llu secondary rays, %.3f ms on the GPU\n", p, s, c, ms);
        bool png = out.size() > 4 && out.substr(out.size() - 4) == ".png";
        bool ok = png ? SaveImage(camera, out) : save_ppm(out, camera->getBitmap(), screen_width, screen_height);
        if (!ok) { fprintf(stderr, "cannot write %s\n", out.c_str()); return 1; }
        fprintf(stderr, "Image saved as: %s\n", out.c_str());
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
