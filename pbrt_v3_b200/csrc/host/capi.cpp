// extern "C" helpers over the host-side scene front end, used by the Python driver (bench.py, tests)
// through ctypes.  These are conveniences around the reference-shaped C++ API (api.h, scene.h): parse
// or synthesise a scene, expose its flattened description, and run the integrator.  The render path
// itself goes through the C ABI in include/pb2.h.
#include "api.h"

using namespace pbrt;

namespace {
std::unique_ptr<FlatScene> g_flat;  // flattened view of the last setup (host memory only)
pb2_camera g_camera;
pb2_film_desc g_film;
pb2_path_params g_params;
std::vector<float> g_rgb;

PathIntegrator *pathIntegrator() {
    RenderSetup *s = pbrtLastSetup();
    return s ? dynamic_cast<PathIntegrator *>(s->integrator.get()) : nullptr;
}
const BVHAccel *sceneBVH() {
    RenderSetup *s = pbrtLastSetup();
    return (s && s->scene) ? dynamic_cast<const BVHAccel *>(s->scene->aggregate.get()) : nullptr;
}
}  // namespace

extern "C" {

int pb2h_error_count(void) { return g_errorCount; }

// Parse a scene without rendering it; returns 0 on success.
int pb2h_parse_file(const char *path, const char *outfile) {
    Options opt;
    if (outfile) opt.imageFile = outfile;
    g_flat.reset();
    if (pbrtIsInitialized()) pbrtCleanup();
    pbrtInit(opt);
    pbrtSetRenderAtWorldEnd(false);
    pbrtParseFile(path);
    return pbrtLastSetup() && pbrtLastSetup()->integrator ? 0 : 1;
}

// The command line's --quick for the next pb2h_parse_* calls (Options::quickRender)
static bool g_nextQuick = false;
void pb2h_set_quick_render(int on) { g_nextQuick = on != 0; }

int pb2h_parse_string(const char *text) {
    Options opt;
    opt.quickRender = g_nextQuick;
    g_flat.reset();
    if (pbrtIsInitialized()) pbrtCleanup();
    pbrtInit(opt);
    pbrtSetRenderAtWorldEnd(false);
    pbrtParseString(text);
    return pbrtLastSetup() && pbrtLastSetup()->integrator ? 0 : 1;
}

void pb2h_cleanup(void) {
    g_flat.reset();
    if (pbrtIsInitialized()) pbrtCleanup();
}

// SURVEY.md §8d synthetic "triangle soup": n_tris random triangles (centre U[-1,1]^3, vertices
// centre + U[-jitter,jitter]^3, pbrt's PCG32 RNG(seed)), matte Kd .6, a ground quad at z=-1.05 and a
// 2-triangle emissive quad (L=40, facing -z) at z=2.5; perspective camera fov 35 from (0,-4.2,.6).
// `header_extra` is optional scene-file text inserted before WorldBegin (e.g. a cropwindow Film).
int pb2h_synth_soup(int64_t n_tris, uint64_t seed, float jitter, int xres, int yres, int spp, int maxdepth,
                    const char *light_strategy) {
    Options opt;
    g_flat.reset();
    if (pbrtIsInitialized()) pbrtCleanup();
    pbrtInit(opt);
    pbrtSetRenderAtWorldEnd(false);
    pbrtLookAt(0, -4.2f, .6f, 0, 0, 0, 0, 0, 1);
    ParamSet cam;
    cam.AddFloat("fov", {35.f});
    pbrtCamera("perspective", cam);
    ParamSet film;
    film.AddInt("xresolution", {xres});
    film.AddInt("yresolution", {yres});
    film.AddString("filename", "soup.pfm");
    pbrtFilm("image", film);
    if (const char *pf = std::getenv("PB2_SOUP_FILTER")) {   // developer switch: the same workload under another reconstruction filter
        ParamSet none;
        pbrtPixelFilter(pf, none);
    }
    if (const char *sm = std::getenv("PB2_SOUP_SPLIT")) {    // developer switch: the same workload over another BVH ("hlbvh")
        ParamSet acc;
        acc.AddString("splitmethod", sm);
        pbrtAccelerator("bvh", acc);
    }
    ParamSet samp;
    samp.AddInt("pixelsamples", {spp});
    pbrtSampler("halton", samp);
    ParamSet integ;
    integ.AddInt("maxdepth", {maxdepth});
    if (light_strategy && *light_strategy) integ.AddString("lightsamplestrategy", light_strategy);
    pbrtIntegrator("path", integ);
    pbrtWorldBegin();
    ParamSet mat;
    mat.AddRGBSpectrum("Kd", {.6f, .6f, .6f});
    pbrtMaterial("matte", mat);
    {
        RNG rng(seed);
        std::vector<Float> P((size_t)9 * n_tris);
        std::vector<int> idx((size_t)3 * n_tris);
        for (int64_t t = 0; t < n_tris; ++t) {
            Float c[3];
            for (int k = 0; k < 3; ++k) c[k] = 2.f * rng.UniformFloat() - 1.f;
            for (int v = 0; v < 3; ++v)
                for (int k = 0; k < 3; ++k) P[9 * t + 3 * v + k] = c[k] + jitter * (2.f * rng.UniformFloat() - 1.f);
            for (int v = 0; v < 3; ++v) idx[3 * t + v] = (int)(3 * t + v);
        }
        ParamSet ps;
        ps.AddPoint3f("P", P);
        ps.AddInt("indices", idx);
        if (n_tris > 0) pbrtShape("trianglemesh", ps);
    }
    {
        ParamSet ps;
        ps.AddPoint3f("P", {-4, -4, -1.05f, 4, -4, -1.05f, 4, 4, -1.05f, -4, 4, -1.05f});
        ps.AddInt("indices", {0, 1, 2, 0, 2, 3});
        pbrtShape("trianglemesh", ps);
    }
    pbrtAttributeBegin();
    {
        ParamSet al;
        al.AddRGBSpectrum("L", {40.f, 40.f, 40.f});
        pbrtAreaLightSource("diffuse", al);
        ParamSet ps;
        ps.AddPoint3f("P", {-1.5f, -1.5f, 2.5f, -1.5f, 1.5f, 2.5f, 1.5f, 1.5f, 2.5f, 1.5f, -1.5f, 2.5f});
        ps.AddInt("indices", {0, 1, 2, 0, 2, 3});
        pbrtShape("trianglemesh", ps);
    }
    pbrtAttributeEnd();
    pbrtWorldEnd();
    return pbrtLastSetup() && pbrtLastSetup()->integrator ? 0 : 1;
}

// Synthetic instanced workload: ONE object - a soup of `n_object_tris` random
// triangles in the unit cube (centre U[-.5,.5]^3, vertices centre + U[-jitter,jitter]^3, RNG(seed)) -
// instanced grid x grid times on a regular grid with a random rotation about z and a small random
// offset per instance (RNG(seed_instances)); matte Kd .6, a floor quad and a 2-triangle area light
// above the grid; perspective camera looking down at the field.
int pb2h_synth_instanced(int64_t n_object_tris, int grid, uint64_t seed, uint64_t seed_instances, float jitter, int xres,
                         int yres, int spp, int maxdepth) {
    Options opt;
    g_flat.reset();
    if (pbrtIsInitialized()) pbrtCleanup();
    pbrtInit(opt);
    pbrtSetRenderAtWorldEnd(false);
    const Float spacing = 1.4f, half = spacing * grid / 2;
    pbrtLookAt(0, -2.2f * half - 2.f, 1.3f * half + 1.5f, 0, 0, 0, 0, 0, 1);
    ParamSet cam;
    cam.AddFloat("fov", {40.f});
    pbrtCamera("perspective", cam);
    ParamSet film;
    film.AddInt("xresolution", {xres});
    film.AddInt("yresolution", {yres});
    film.AddString("filename", "instanced.pfm");
    pbrtFilm("image", film);
    ParamSet samp;
    samp.AddInt("pixelsamples", {spp});
    pbrtSampler("halton", samp);
    ParamSet integ;
    integ.AddInt("maxdepth", {maxdepth});
    pbrtIntegrator("path", integ);
    pbrtWorldBegin();
    ParamSet mat;
    mat.AddRGBSpectrum("Kd", {.6f, .6f, .6f});
    pbrtMaterial("matte", mat);
    pbrtObjectBegin("soup");
    {
        RNG rng(seed);
        std::vector<Float> P((size_t)9 * n_object_tris);
        std::vector<int> idx((size_t)3 * n_object_tris);
        for (int64_t t = 0; t < n_object_tris; ++t) {
            Float c[3];
            for (int k = 0; k < 3; ++k) c[k] = rng.UniformFloat() - .5f;
            for (int v = 0; v < 3; ++v)
                for (int k = 0; k < 3; ++k) P[9 * t + 3 * v + k] = c[k] + jitter * (2.f * rng.UniformFloat() - 1.f);
            for (int v = 0; v < 3; ++v) idx[3 * t + v] = (int)(3 * t + v);
        }
        ParamSet ps;
        ps.AddPoint3f("P", P);
        ps.AddInt("indices", idx);
        if (n_object_tris > 0) pbrtShape("trianglemesh", ps);
    }
    pbrtObjectEnd();
    {
        RNG rng(seed_instances);
        for (int gy = 0; gy < grid; ++gy)
            for (int gx = 0; gx < grid; ++gx) {
                pbrtAttributeBegin();
                Float tx = (gx + .5f) * spacing - half + .2f * (rng.UniformFloat() - .5f);
                Float ty = (gy + .5f) * spacing - half + .2f * (rng.UniformFloat() - .5f);
                pbrtTranslate(tx, ty, .55f);
                pbrtRotate(360.f * rng.UniformFloat(), 0, 0, 1);
                pbrtObjectInstance("soup");
                pbrtAttributeEnd();
            }
    }
    {
        ParamSet ps;
        Float e = half + 2.f;
        ps.AddPoint3f("P", {-e, -e, 0, e, -e, 0, e, e, 0, -e, e, 0});
        ps.AddInt("indices", {0, 1, 2, 0, 2, 3});
        pbrtShape("trianglemesh", ps);
    }
    pbrtAttributeBegin();
    {
        ParamSet al;
        al.AddRGBSpectrum("L", {30.f, 30.f, 30.f});
        pbrtAreaLightSource("diffuse", al);
        ParamSet ps;
        Float e = .6f * half + .5f, z = half + 3.f;
        ps.AddPoint3f("P", {-e, -e, z, -e, e, z, e, e, z, e, -e, z});
        ps.AddInt("indices", {0, 1, 2, 0, 2, 3});
        pbrtShape("trianglemesh", ps);
    }
    pbrtAttributeEnd();
    pbrtWorldEnd();
    return pbrtLastSetup() && pbrtLastSetup()->integrator ? 0 : 1;
}

// Flattened description of the last parsed scene (host memory, valid until the next parse/cleanup).
const pb2_scene_desc *pb2h_scene_desc(void) {
    if (g_flat) return &g_flat->desc;
    const BVHAccel *bvh = sceneBVH();
    PathIntegrator *pi = pathIntegrator();
    if (!bvh || !pi) return nullptr;
    RenderSetup *s = pbrtLastSetup();
    g_flat = FlattenScene(*bvh, s->scene->lights, pi->LightSampleStrategy());
    return g_flat ? &g_flat->desc : nullptr;
}
void pb2h_set_light_strategy(int strategy) {
    if (pb2h_scene_desc()) g_flat->desc.light_strategy = strategy;
}
const pb2_camera *pb2h_camera(void) {
    RenderSetup *s = pbrtLastSetup();
    if (!s || !s->camera) return nullptr;
    g_camera = s->camera->Desc();
    return &g_camera;
}
const pb2_film_desc *pb2h_film(void) {
    RenderSetup *s = pbrtLastSetup();
    if (!s || !s->film) return nullptr;
    g_film = s->film->Desc();
    return &g_film;
}
const pb2_path_params *pb2h_path_params(void) {
    PathIntegrator *pi = pathIntegrator();
    if (!pi) return nullptr;
    g_params = pi->Params();
    return &g_params;
}

// Integrator::Render on the device; the resolved RGB image (film.cpp:178-203) is kept and can be
// fetched with pb2h_image().  Returns 0 on success.
int pb2h_render(int write_image, pb2_stats *stats) {
    RenderSetup *s = pbrtLastSetup();
    PathIntegrator *pi = pathIntegrator();
    if (!s || !pi) return 1;
    int before = g_errorCount;
    for (auto &px : s->film->pixels) px = Film::Pixel();
    pi->writeImage = write_image != 0;
    pi->Render(*s->scene);
    if (stats) *stats = pi->lastStats;
    if (g_errorCount != before) return 2;
    g_rgb = s->film->ResolveRGB();
    return 0;
}
const float *pb2h_image(int *width, int *height) {
    RenderSetup *s = pbrtLastSetup();
    if (!s || !s->film) return nullptr;
    *width = s->film->croppedPixelBounds.pMax.x - s->film->croppedPixelBounds.pMin.x;
    *height = s->film->croppedPixelBounds.pMax.y - s->film->croppedPixelBounds.pMin.y;
    return g_rgb.data();
}
// Film::MergeFilmTile + WriteImage arithmetic applied to an externally produced (e.g. NCCL-reduced)
// rgbw buffer: out_rgb gets 3 floats per pixel.
int pb2h_resolve_film(const float *rgbw, float *out_rgb) {
    RenderSetup *s = pbrtLastSetup();
    if (!s || !s->film) return 1;
    for (auto &px : s->film->pixels) px = Film::Pixel();
    s->film->MergeDeviceFilm(rgbw);
    std::vector<Float> rgb = s->film->ResolveRGB();
    std::memcpy(out_rgb, rgb.data(), rgb.size() * sizeof(float));
    return 0;
}
// Device handle of the last scene (created on first use); NULL without a CUDA device.
pb2_scene *pb2h_device_scene(void) {
    const BVHAccel *bvh = sceneBVH();
    RenderSetup *s = pbrtLastSetup();
    PathIntegrator *pi = pathIntegrator();
    if (!bvh || !s || !pi) return nullptr;
    std::shared_ptr<DeviceScene> ds = GetDeviceScene(*bvh, s->scene->lights, pi->LightSampleStrategy());
    return ds ? DeviceSceneHandle(*ds) : nullptr;
}

// Scene::Intersect through the host classes (scene.cpp:45-49 -> BVHAccel::Intersect -> the device), as a caller of the
// reference's API would issue it before or without Integrator::Render.  Returns 1 on a hit and fills t.
int pb2h_scene_intersect(const float *o, const float *d, float *t_hit) {
    RenderSetup *s = pbrtLastSetup();
    if (!s || !s->scene) return -1;
    Ray ray(Point3f(o[0], o[1], o[2]), Vector3f(d[0], d[1], d[2]));
    SurfaceInteraction isect;
    if (!s->scene->Intersect(ray, &isect)) return 0;
    if (t_hit) *t_hit = ray.tMax;
    return 1;
}

int pb2h_write_pfm(const char *path, const float *rgb, int w, int h) { return WriteImagePFM(path, rgb, w, h) ? 0 : 1; }
// WriteImage (imageio.cpp:81-122): EXR / PFM / PNG / TGA by extension; the window arguments only matter for EXR
// ReadImage (imageio.cpp:60-79): PFM / PNG / TGA / OpenEXR by extension; rgb (3 floats per pixel, row 0 at the top) may be
// NULL to ask for the resolution only
int pb2h_read_image(const char *path, float *rgb, int *w, int *h) {
    std::vector<float> data;
    if (!ReadImage(path, &data, w, h)) return 1;
    if (rgb) std::memcpy(rgb, data.data(), data.size() * sizeof(float));
    return 0;
}
int pb2h_write_image(const char *path, const float *rgb, int w, int h, int total_w, int total_h, int x_offset, int y_offset) {
    return WriteImage(path, rgb, w, h, total_w, total_h, x_offset, y_offset) ? 0 : 1;
}

// Loop subdivision of a control mesh (tests compare with the reference's CreateLoopSubdiv).
// Call once with out pointers NULL to get the sizes, then again with buffers.
int pb2h_loop_subdivide(int n_levels, int n_indices, const int *indices, int n_vertices, const float *P,
                        int *out_n_vertices, int *out_n_indices, float *out_P, float *out_N, int *out_indices) {
    std::vector<Point3f> pin(n_vertices), pl;
    for (int i = 0; i < n_vertices; ++i) pin[i] = Point3f(P[3 * i], P[3 * i + 1], P[3 * i + 2]);
    std::vector<Normal3f> ns;
    std::vector<int> idx;
    LoopSubdivide(n_levels, n_indices, indices, n_vertices, pin.data(), &pl, &ns, &idx);
    *out_n_vertices = (int)pl.size();
    *out_n_indices = (int)idx.size();
    if (out_P)
        for (size_t i = 0; i < pl.size(); ++i) { out_P[3 * i] = pl[i].x; out_P[3 * i + 1] = pl[i].y; out_P[3 * i + 2] = pl[i].z; }
    if (out_N)
        for (size_t i = 0; i < ns.size(); ++i) { out_N[3 * i] = ns[i].x; out_N[3 * i + 1] = ns[i].y; out_N[3 * i + 2] = ns[i].z; }
    if (out_indices) std::memcpy(out_indices, idx.data(), idx.size() * sizeof(int));
    return 0;
}

}  // extern "C"
