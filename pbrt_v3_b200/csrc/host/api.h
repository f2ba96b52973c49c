// The scene-description API of the reference (src/core/api.h:47-117), restricted to the directives
// the path-tracing configurations use.  Same function names and argument meaning; the parser
// (parser.cpp) and programmatic callers drive these exactly as they drive the reference.
#ifndef PB2_HOST_API_H
#define PB2_HOST_API_H

#include "scene.h"

namespace pbrt {

struct Options {
    int nThreads = 0;  // accepted for CLI compatibility; the GPU path does not use host threads
    bool quickRender = false;  // --quick (pbrt.cpp:116-117): quarter resolution, one sample per pixel
    bool quiet = false;
    std::string imageFile;
    Float cropWindow[2][2] = {{0, 1}, {0, 1}};
};

// -- life cycle (api.cpp:682-722)
void pbrtInit(const Options &options);
bool pbrtIsInitialized();
void pbrtCleanup();

// -- current transformation matrix (api.cpp:724-872).  Matrices arrive column-major, as in scene files.
void pbrtIdentity();
void pbrtTranslate(Float tx, Float ty, Float tz);
void pbrtRotate(Float degrees, Float axisX, Float axisY, Float axisZ);
void pbrtScale(Float x, Float y, Float z);
void pbrtLookAt(Float eyeX, Float eyeY, Float eyeZ, Float lookX, Float lookY, Float lookZ, Float upX, Float upY, Float upZ);
void pbrtConcatTransform(Float columnMajor[16]);
void pbrtTransform(Float columnMajor[16]);
void pbrtCoordinateSystem(const std::string &coordSysName), pbrtCoordSysTransform(const std::string &coordSysName);

// -- options block: the one-per-scene plugins (api.cpp:874-1010)
typedef void PluginDirective(const std::string &pluginName, const ParamSet &pluginParams);
PluginDirective pbrtPixelFilter, pbrtFilm, pbrtSampler, pbrtAccelerator, pbrtIntegrator, pbrtCamera;

// -- world block: attribute / transform stacks, materials, lights, shapes, objects (api.cpp:1023-1588)
void pbrtWorldBegin(), pbrtWorldEnd();
void pbrtAttributeBegin(), pbrtAttributeEnd(), pbrtTransformBegin(), pbrtTransformEnd();
PluginDirective pbrtMaterial, pbrtMakeNamedMaterial, pbrtLightSource, pbrtAreaLightSource, pbrtShape;
void pbrtNamedMaterial(const std::string &materialName);
void pbrtTexture(const std::string &name, const std::string &type, const std::string &texname, const ParamSet &params);
void pbrtObjectBegin(const std::string &objectName), pbrtObjectEnd(), pbrtObjectInstance(const std::string &objectName);
void pbrtReverseOrientation();

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
