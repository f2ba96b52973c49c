// The scene-description API of the reference (src/core/api.h:47-117), restricted to the directives
// the path-tracing configurations use.  Same function names and argument meaning; the parser
// (parser.cpp) and programmatic callers drive these exactly as they drive the reference.
#ifndef PB2_HOST_API_H
#define PB2_HOST_API_H

#include "scene.h"

namespace pbrt {

struct Options {
    int nThreads = 0;  // accepted for CLI compatibility; the GPU path does not use host threads
    bool quiet = false;
    std::string imageFile;
    Float cropWindow[2][2] = {{0, 1}, {0, 1}};
};

void pbrtInit(const Options &opt);
bool pbrtIsInitialized();
void pbrtCleanup();
void pbrtIdentity();
void pbrtTranslate(Float dx, Float dy, Float dz);
void pbrtRotate(Float angle, Float ax, Float ay, Float az);
void pbrtScale(Float sx, Float sy, Float sz);
void pbrtLookAt(Float ex, Float ey, Float ez, Float lx, Float ly, Float lz, Float ux, Float uy, Float uz);
void pbrtConcatTransform(Float transform[16]);
void pbrtTransform(Float transform[16]);
void pbrtCoordinateSystem(const std::string &);
void pbrtCoordSysTransform(const std::string &);
void pbrtPixelFilter(const std::string &name, const ParamSet &params);
void pbrtFilm(const std::string &type, const ParamSet &params);
void pbrtSampler(const std::string &name, const ParamSet &params);
void pbrtAccelerator(const std::string &name, const ParamSet &params);
void pbrtIntegrator(const std::string &name, const ParamSet &params);
void pbrtCamera(const std::string &, const ParamSet &cameraParams);
void pbrtWorldBegin();
void pbrtAttributeBegin();
void pbrtAttributeEnd();
void pbrtTransformBegin();
void pbrtTransformEnd();
void pbrtMaterial(const std::string &name, const ParamSet &params);
void pbrtMakeNamedMaterial(const std::string &name, const ParamSet &params);
void pbrtNamedMaterial(const std::string &name);
void pbrtAreaLightSource(const std::string &name, const ParamSet &params);
void pbrtShape(const std::string &name, const ParamSet &params);
void pbrtObjectBegin(const std::string &name);
void pbrtObjectEnd();
void pbrtObjectInstance(const std::string &name);
void pbrtReverseOrientation();
void pbrtWorldEnd();

void pbrtParseFile(std::string filename);
void pbrtParseString(std::string str);

// What pbrtWorldEnd() built, kept alive until pbrtCleanup() so that embedding code (tests, bench.py
// through capi.cpp) can re-render, query intersections or read the film without re-parsing.
struct RenderSetup {
    std::unique_ptr<Scene> scene;
    std::shared_ptr<Camera> camera;
    std::shared_ptr<Sampler> sampler;
    std::unique_ptr<Integrator> integrator;
    std::unique_ptr<Film> film;
};
RenderSetup *pbrtLastSetup();
// When false, pbrtWorldEnd() builds the RenderSetup but does not call Integrator::Render.
void pbrtSetRenderAtWorldEnd(bool render);

}  // namespace pbrt
#endif
