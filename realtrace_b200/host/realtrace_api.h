// realtrace_api.h — C++ mirror of the reference's Serial scene/render API.
//
// Same class names, constructor signatures, public fields and method names as
// /root/reference/Serial/{vector3D,color,ray,material,object,sphere,plane,cylinder,triangle,
// lightsource,pointlightsource,world,camera,renderengine}.h, so host code written against the
// reference (Serial/lumina.cpp:302-370) compiles against this header unchanged.  The per-file
// headers of the reference exist next to this file as one-line forwarders.
//
// What differs: nothing here traces rays on the CPU.  RenderEngine::render()/renderLoop(),
// World::firstIntersection() and World::shade_ray() flatten the scene once per revision and call
// the C ABI of include/realtrace_b200.h (LBVH build + wavefront kernels on the GPU); without a
// CUDA device they throw std::runtime_error — there is no CPU fallback.
// Additions the reference lacks: RenderEngine::render(), setMaxDepth(), frameStats(); read
// accessors on Sphere/Plane/Cylinder/Triangle (their fields are private there); World::invalidate().
#ifndef REALTRACE_API_H
#define REALTRACE_API_H

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <iostream>
#include <string>
#include <vector>

struct rt_ctx;
struct rt_frame_stats;

// ---- vector3D.h:7-56 -----------------------------------------------------------------------------
class Vector3D {
public:
    double e[3];
    Vector3D() { e[0] = e[1] = e[2] = 0.0; }
    Vector3D(double e0, double e1, double e2) { e[0] = e0; e[1] = e1; e[2] = e2; }
    double X() const { return e[0]; }
    double Y() const { return e[1]; }
    double Z() const { return e[2]; }
    void X(double x) { e[0] = x; }
    void Y(double y) { e[1] = y; }
    void Z(double z) { e[2] = z; }
    const Vector3D& operator+() const { return *this; }
    Vector3D operator-() const { return Vector3D(-e[0], -e[1], -e[2]); }
    double operator[](int i) const { return e[i]; }
    double& operator[](int i) { return e[i]; }
    Vector3D& operator+=(const Vector3D& v) { for (int k = 0; k < 3; k++) e[k] += v.e[k]; return *this; }
    Vector3D& operator-=(const Vector3D& v) { for (int k = 0; k < 3; k++) e[k] -= v.e[k]; return *this; }
    Vector3D& operator*=(double s) { for (int k = 0; k < 3; k++) e[k] *= s; return *this; }
    Vector3D& operator/=(double s) { for (int k = 0; k < 3; k++) e[k] /= s; return *this; }
    double squaredlength() const { return e[0] * e[0] + e[1] * e[1] + e[2] * e[2]; }
    double length() const { return std::sqrt(squaredlength()); }
    void normalize() { double l = length(); for (int k = 0; k < 3; k++) e[k] /= l; }
};
inline bool operator==(const Vector3D& a, const Vector3D& b) { return a.e[0] == b.e[0] && a.e[1] == b.e[1] && a.e[2] == b.e[2]; }
inline bool operator!=(const Vector3D& a, const Vector3D& b) { return !(a == b); }
inline Vector3D operator+(const Vector3D& a, const Vector3D& b) { return Vector3D(a.e[0] + b.e[0], a.e[1] + b.e[1], a.e[2] + b.e[2]); }
inline Vector3D operator-(const Vector3D& a, const Vector3D& b) { return Vector3D(a.e[0] - b.e[0], a.e[1] - b.e[1], a.e[2] - b.e[2]); }
inline Vector3D operator/(const Vector3D& a, double s) { return Vector3D(a.e[0] / s, a.e[1] / s, a.e[2] / s); }
inline Vector3D operator*(const Vector3D& a, double s) { return Vector3D(a.e[0] * s, a.e[1] * s, a.e[2] * s); }
inline Vector3D operator*(double s, const Vector3D& a) { return a * s; }
inline Vector3D operator*(const Vector3D& a, const Vector3D& b) { return Vector3D(a.e[0] * b.e[0], a.e[1] * b.e[1], a.e[2] * b.e[2]); }
inline Vector3D unitVector(const Vector3D& v) { return v / v.length(); }
inline Vector3D crossProduct(const Vector3D& a, const Vector3D& b) {
    return Vector3D(a.Y() * b.Z() - a.Z() * b.Y(), a.Z() * b.X() - a.X() * b.Z(), a.X() * b.Y() - a.Y() * b.X());
}
inline double dotProduct(const Vector3D& a, const Vector3D& b) { return a.X() * b.X() + a.Y() * b.Y() + a.Z() * b.Z(); }
inline double tripleProduct(const Vector3D& a, const Vector3D& b, const Vector3D& c) { return dotProduct(crossProduct(a, b), c); }
inline std::ostream& operator<<(std::ostream& o, const Vector3D& v) { return o << "{" << v.X() << "," << v.Y() << "," << v.Z() << "}"; }

// ---- color.h:5-35 --------------------------------------------------------------------------------
class Color {
public:
    double r, g, b;
    Color(double val = 0.0) { r = g = b = val; }
    Color(double red, double green, double blue) : r(red), g(green), b(blue) {}
    Color(const Vector3D& v) : r(v.X()), g(v.Y()), b(v.Z()) {}
    void R(double v) { r = v; }
    void G(double v) { g = v; }
    void B(double v) { b = v; }
    double R() const { return r; }
    double G() const { return g; }
    double B() const { return b; }
    void clamp() {
        r = r > 1.0 ? 1.0 : (r < 0.0 ? 0.0 : r);           // color.cpp:19-28
        g = g > 1.0 ? 1.0 : (g < 0.0 ? 0.0 : g);
        b = b > 1.0 ? 1.0 : (b < 0.0 ? 0.0 : b);
    }
};
inline Color operator*(const Color& c, double f) { return Color(c.r * f, c.g * f, c.b * f); }
inline Color operator*(double f, const Color& c) { return c * f; }
inline Color operator*(const Color& a, const Color& b) { return Color(a.r * b.r, a.g * b.g, a.b * b.b); }
inline Color operator/(const Color& c, double f) { return Color(c.r / f, c.g / f, c.b / f); }
inline Color operator+(const Color& a, const Color& b) { return Color(a.r + b.r, a.g + b.g, a.b + b.b); }

// ---- ray.h:10-47 ---------------------------------------------------------------------------------
class Object;
const float SMALLEST_DIST = 1e-4f;
class Ray {
    Vector3D origin, direction;
    float t;
    bool hit;
    int idx;
    const Object* object;
    int level;
    float refractive_index;
    Vector3D normal;
public:
    Ray(const Vector3D& o, const Vector3D& d, int _level = 0, float _ref_idx = 1.0f)
        : origin(o), direction(d), t(FLT_MAX), hit(false), idx(-1), object(nullptr), level(_level), refractive_index(_ref_idx) {
        direction.normalize();
    }
    Vector3D getOrigin() const { return origin; }
    Vector3D getDirection() const { return direction; }
    Vector3D getPosition() const { return origin + (double)t * direction; }
    Vector3D getNormal() const { return normal; }
    void setNormal(const Vector3D n) { normal = n; }
    float getParameter() const { return t; }
    void strictSetParameter(const float par) { t = par; }
    bool setParameter(const float par, const Object* obj) {            // ray.cpp:3-13
        if (par < t && par > SMALLEST_DIST) { hit = true; t = par; object = obj; return true; }
        return false;
    }
    bool didHit() const { return hit; }
    void setHit(bool f) { hit = f; }
    void setIdx(int i) { idx = i; }
    int getIdx() { return idx; }
    const Object* intersected() const { return object; }
    int getLevel() const { return level; }
};

// ---- material.h:11-50 ----------------------------------------------------------------------------
class World;
class Material {
protected:
    World* world;
public:
    Color color;
    double ka, kd, ks, kr, kt, eta, n;
    Material(World* w) : world(w), color(0), ka(0.2), kd(1.0), ks(0.4), kr(0), kt(0), eta(128), n(0) {}
    virtual ~Material() {}
    virtual Color shade(const Ray& incident, const bool isSolid = true) const;
    virtual bool isBarycentric() const { return false; }
};
class BarycentricMaterial : public Material {
protected:
    Vector3D vertexA, vertexB, vertexC;
    std::vector<Color> colors;
public:
    BarycentricMaterial(World* w, const Vector3D v1, const Vector3D v2, const Vector3D v3, const Color& c1,
                        const Color& c2, const Color& c3)
        : Material(w), vertexA(v1), vertexB(v2), vertexC(v3) { colors.push_back(c1); colors.push_back(c2); colors.push_back(c3); }
    Color shade(const Ray& incident, const bool isSolid = true) const override;
    bool isBarycentric() const override { return true; }
    const Color& vertexColor(int i) const { return colors[i]; }
};

// ---- utilities.h:10-18 (the reference's axis_max = numeric_limits::min() quirk is not kept) -------
class BBox {
public:
    double axis_min[3], axis_max[3];
    BBox() { for (int k = 0; k < 3; k++) { axis_min[k] = DBL_MAX; axis_max[k] = -DBL_MAX; } }
};

// ---- object.h:10-24 ------------------------------------------------------------------------------
class Object {
protected:
    Material* material;
    bool isSolid;
public:
    enum Kind { SPHERE, PLANE, CYLINDER, TRIANGLE };
    Object(Material* mat) : material(mat), isSolid(true) {}
    virtual ~Object() {}
    // object.h:15: one ray against THIS object alone, with the reference's acceptance rules
    // (Ray::setParameter).  A host-side convenience of the class mirror (plain FP64 like the
    // reference); rendering and World::firstIntersection never call it — they run on the GPU.
    virtual bool intersect(Ray& ray) const = 0;
    virtual Color shade(const Ray& ray) const { return material->shade(ray, isSolid); }
    Material* getMaterial() const { return material; }
    virtual Vector3D getNormalAtPosition(const Vector3D& position) const = 0;
    virtual Kind kind() const = 0;
};

class Sphere : public Object {                              // sphere.h:10-25
    Vector3D position;
    double radius;
public:
    Sphere(const Vector3D& _pos, double _rad, Material* mat) : Object(mat), position(_pos), radius(_rad) {}
    bool intersect(Ray& ray) const override;               // sphere.cpp:5-39
    Vector3D getNormalAtPosition(const Vector3D& p) const override { return p - position; }
    Kind kind() const override { return SPHERE; }
    const Vector3D& getPosition() const { return position; }
    double getRadius() const { return radius; }
};

class Plane : public Object {                               // plane.h:10-30
    Vector3D position1, position2, position3, position4, Normal;
public:
    Plane(const Vector3D& p1, const Vector3D& p2, const Vector3D& p3, const Vector3D& p4, Material* mat)
        : Object(mat), position1(p1), position2(p2), position3(p3), position4(p4) { Normal = crossProduct(p3 - p1, p2 - p1); }
    bool intersect(Ray& ray) const override;               // plane.cpp:12-27
    Vector3D getNormalAtPosition(const Vector3D&) const override { return Normal; }
    Kind kind() const override { return PLANE; }
    const Vector3D& getCorner(int i) const { return i == 0 ? position1 : i == 1 ? position2 : i == 2 ? position3 : position4; }
};

class Cylinder : public Object {                            // cylinder.h:10-25
    Vector3D position;
    double radius;
    Vector3D up;
public:
    Cylinder(const Vector3D& _pos, double _rad, const Vector3D& u, Material* mat) : Object(mat), position(_pos), radius(_rad), up(u) {}
    bool intersect(Ray& ray) const override;               // cylinder.cpp:14-32
    Vector3D getNormalAtPosition(const Vector3D& p) const override {
        double t = dotProduct(p - position, up) / dotProduct(up, up);
        return p - position - t * up;
    }
    Kind kind() const override { return CYLINDER; }
    const Vector3D& getPosition() const { return position; }
    double getRadius() const { return radius; }
    const Vector3D& getUp() const { return up; }
};

class Triangle : public Object {                            // triangle.h:14-37
    Vector3D vertexA, vertexB, vertexC;
public:
    Triangle(const Vector3D& a, const Vector3D& b, const Vector3D& c, Material* mat) : Object(mat), vertexA(a), vertexB(b), vertexC(c) {}
    bool intersect(Ray& ray) const override;               // triangle.cpp:10-24
    Vector3D getNormalAtPosition(const Vector3D&) const override { return crossProduct(vertexA - vertexB, vertexA - vertexC); }
    Kind kind() const override { return TRIANGLE; }
    BBox getWorldBound();
    Vector3D getVertex(int i) const { return i == 0 ? vertexA : i == 1 ? vertexB : vertexC; }
};

// ---- lightsource.h:8-19, pointlightsource.h:6-14 -------------------------------------------------
class LightSource {
protected:
    World* world;
    Color intensity;
public:
    LightSource(World* _world, const Color _intensity) : world(_world), intensity(_intensity) {}
    virtual ~LightSource() {}
    virtual Vector3D getPosition() const = 0;
    Color getIntensity() const { return intensity; }
};
class PointLightSource : public LightSource {
    Vector3D position;
public:
    PointLightSource(World* _world, const Vector3D& _pos, const Color& _intensity) : LightSource(_world, _intensity), position(_pos) {}
    Vector3D getPosition() const override { return position; }
};

// ---- uniform-grid.h:22-49: kept only so that `world->uniform_grid = UniformGrid(all_triangles)`
// (lumina.cpp:289) still compiles; the LBVH is built on the GPU at the next render.
class UniformGrid {
public:
    UniformGrid() {}
    UniformGrid(std::vector<Triangle*>&) {}
};

// ---- world.h:12-45 -------------------------------------------------------------------------------
namespace rtb200 { struct Device; }
class World {
    std::vector<Object*> objectList;
    std::vector<LightSource*> lightSourceList;
    Color ambient, background;
    unsigned long revision = 1;
    rtb200::Device* dev = nullptr;
    friend class RenderEngine;
    friend struct rtb200::Device;
public:
    UniformGrid uniform_grid;
    World() : objectList(0), lightSourceList(0), ambient(0), background(0) {}
    ~World();
    World(const World&) = delete;
    World& operator=(const World&) = delete;
    void setBackground(const Color& bk) { background = bk; revision++; }
    Color getbackground() { return background; }
    void setAmbient(const Color& amb) { ambient = amb; revision++; }
    Color getAmbient() { return ambient; }
    void addLight(LightSource* ls) { lightSourceList.push_back(ls); revision++; }
    void addObject(Object* obj) { objectList.push_back(obj); revision++; }
    std::vector<Object*>& getObjectList() { return objectList; }
    // Call after editing an object or material in place (the reference's public fields can be
    // changed at any time; geometry and materials are re-flattened on the next render).
    void invalidate() { revision++; }
    void setMaxDepth(int d) { max_depth = d; }             // replaces #define RECURSION_DEPTH (world.h:11)
    int getMaxDepth() const { return max_depth; }
    float firstIntersection(Ray& ray);                      // world.cpp:5-17, one-ray GPU query
    Color shade_ray(Ray ray);                               // world.cpp:32-111, one-ray GPU query
private:
    int max_depth = 10;
};

// ---- camera.h:7-34 -------------------------------------------------------------------------------
class Camera {
    Vector3D position, target, up, line_of_sight, u, v, w;
    unsigned char* bitmap;                                  // page-locked (rt_host_alloc) where a CUDA device exists,
    bool bitmap_pinned;                                     // so the frame arrives by asynchronous DMA
    int width, height;
    float fovy, focalDistance, focalWidth, focalHeight, aspect;
    friend class RenderEngine;
public:
    Camera(const Vector3D& _pos, const Vector3D& _target, const Vector3D& _up, float fovy, int w, int h);
    ~Camera();
    Camera(const Camera&) = delete;
    Camera& operator=(const Camera&) = delete;
    const Vector3D get_ray_direction(const int i, const int j) const;
    const Vector3D& get_position() const { return position; }
    void drawPixel(int i, int j, Color c);
    unsigned char* getBitmap() { return bitmap; }
    int getWidth() { return width; }
    int getHeight() { return height; }
};

// ---- renderengine.h:7-18 -------------------------------------------------------------------------
class RenderEngine {
    World* world;
    Camera* camera;
public:
    RenderEngine(World* _world, Camera* _camera) : world(_world), camera(_camera) {}
    // The reference renders one column per call and returns true after the last one
    // (renderengine.cpp:10-26).  Here one call renders the whole frame on the GPU into
    // Camera::getBitmap() and returns true, so `while(!engine->renderLoop()){}` still terminates.
    bool renderLoop() { render(); return true; }
    void render();                                          // the name BASELINE.json's north star uses; the frame
                                                            // is split over every visible GPU (setDevices narrows it)
    // Which CUDA devices World's GPU core uses (default: all visible; call before the first render).
    static void setDevices(const std::vector<int>& device_ids);
    void setMaxDepth(int d) { world->setMaxDepth(d); }
    // rays / timings of the last frame (include/realtrace_b200.h: rt_frame_stats)
    void frameStats(unsigned long long& primary, unsigned long long& shadow, unsigned long long& secondary, float& ms_device) const;
};

// ---- lumina.cpp:195-290: OBJ loader (SCALING_FACTOR 15).  The same call loads the same scene as the
// reference: at most 2000 faces (lumina.cpp:266) unless max_faces says otherwise (< 0: no cap).
// texel_mode selects how a texture is fetched (lumina.cpp:175-187 goes through the un-vendored DevIL, see
// DESIGN.md section 2): LITERAL restates those lines on an RGBA8 top-left-origin image (swapped row/column,
// no /255, the un-decremented vt index of :249); NORMALISED is texel/255 with the usual OBJ conventions.
enum { RT_TEXEL_NORMALISED = 0, RT_TEXEL_LITERAL = 1 };
const int LUMINA_MAX_FACES = 2000;
void load_image_from_obj(World* world, std::string file_name, std::string texture_file_name = "",
                         std::string occlusion_map_file_name = "", int max_faces = LUMINA_MAX_FACES,
                         int texel_mode = RT_TEXEL_NORMALISED);
void init_material_from_obj(Material* m);                   // lumina.cpp:163-172

// ---- lumina.cpp:424-439 SaveImage: the frame as an image file.  The reference goes through DevIL
// (ilTexImage + ilSave(IL_PNG)); here a self-contained writer (zlib deflate, 8-bit RGB, no interlace).
// Row 0 of Camera::getBitmap() is the BOTTOM row on screen (vshader.vs:8), so rows are written top-down.
bool save_png(const std::string& file_name, const unsigned char* rgb_bottom_up, int width, int height);
bool save_ppm(const std::string& file_name, const unsigned char* rgb_bottom_up, int width, int height);
inline bool SaveImage(Camera* camera, const std::string& file_name) {
    return save_png(file_name, camera->getBitmap(), camera->getWidth(), camera->getHeight());
}

// ---- Parellel/interactive_camera.cu:7-102: the orbit camera of the reference's CUDA app (BASELINE
// config 5).  Same fields, clamps and eye-position formula (float arithmetic, :64-72); makeCamera()
// returns a Serial-style Camera looking at centerPosition.
class InteractiveCamera {
public:
    float centerPosition[3];
    float yaw, pitch, radius, apertureRadius;
    float resolution[2], fov[2];
    InteractiveCamera();
    void changeYaw(float m);
    void changePitch(float m);
    void changeRadius(float m);
    void changeAltitude(float m);
    void setResolution(float x, float y);
    void setFOVX(float fovx);
    void eyePosition(float out[3]) const;                   // buildRenderCamera, :64-70
    Camera* makeCamera() const;                             // caller owns the result
};

#endif
