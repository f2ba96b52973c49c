// TEST INFRASTRUCTURE ONLY.  The binding of INTEGRATION.md section 2-3, COMPILED against the reference's own classes: a
// `B200PathIntegrator : pbrt::Integrator` whose Render(const Scene &) flattens the REFERENCE's Scene / BVHAccel / Triangle /
// Sphere / materials / lights / Film / PerspectiveCamera objects into a pb2_scene_desc and renders through libpb2.so, then
// hands the merged film back to the reference's Film::MergeFilmTile and WriteImage.  Built only where /root/reference
// exists (oracle/Makefile.ref -> oracle/_ref/libb200_overlay.so, linked against pbrt_v3_b200/lib/libpb2.so) and exercised
// by tests/test_gpu_overlay.py: the same reference Scene rendered by the reference's PathIntegrator and by this integrator
// must agree - which proves the drop-in claim with the reference's classes in the loop rather than our mirrored ones.
//
// A maintainer would add `friend class B200PathIntegrator;` to BVHAccel, Triangle, Sphere, the materials and
// DiffuseAreaLight (or accessors); this file sees their private members the way the harness does.
#include "ref_harness.cpp"   // the harness (unity build): RefScene, makeRenderObjects, the `#define private public` includes

namespace pbrt {

class B200PathIntegrator : public Integrator {
  public:
    B200PathIntegrator(int maxDepth, std::shared_ptr<const Camera> camera, std::shared_ptr<Sampler> sampler, const Bounds2i &pixelBounds,
                       Float rrThreshold, const std::string &lightStrategy)
        : maxDepth(maxDepth), camera(camera), sampler(sampler), pixelBounds(pixelBounds), rrThreshold(rrThreshold), lightStrategy(lightStrategy) {}
    void Render(const Scene &scene) override;
    pb2_stats stats{};
    std::string error;

  private:
    const int maxDepth;
    std::shared_ptr<const Camera> camera;
    std::shared_ptr<Sampler> sampler;
    const Bounds2i pixelBounds;
    const Float rrThreshold;
    const std::string lightStrategy;
};

namespace {

// Everything a pb2_scene_desc points to, owned here for the duration of pb2_scene_create.
struct Flat {
    std::vector<float> P, N, UV, S;
    std::vector<int32_t> triIndex, triMesh, primIndex, primMaterial, primLight, bvhPrims;
    std::vector<uint8_t> primType;
    std::vector<pb2_mesh> meshes;
    std::vector<pb2_sphere> spheres;
    std::vector<pb2_material> materials;
    std::vector<pb2_light> lights;
    std::vector<pb2_bvh_node> nodes;
    pb2_scene_desc desc;
};

void copyMatrix(const Matrix4x4 &m, float *out) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) out[4 * i + j] = m.m[i][j];
}

template <typename T>
T constantValue(const std::shared_ptr<Texture<T>> &tex, bool *ok) {
    // constant textures are evaluated once; anything that varies over the surface is refused
    if (!dynamic_cast<const ConstantTexture<T> *>(tex.get())) *ok = false;
    SurfaceInteraction si;
    return tex->Evaluate(si);
}

bool FlattenForPb2(const Scene &scene, const std::string &strategy, Flat *f, std::string *err) {
    const BVHAccel *bvh = dynamic_cast<const BVHAccel *>(scene.aggregate.get());
    if (!bvh || !bvh->nodes) {
        *err = "the aggregate is not a BVHAccel";
        return false;
    }
    // LinearBVHNode[] verbatim (same 32-byte layout); the node count is not stored by BVHAccel: walk the tree once
    const LinearBVHNode *ln = reinterpret_cast<const LinearBVHNode *>(bvh->nodes);
    int64_t nNodes = 0;
    {
        std::vector<int> todo{0};
        while (!todo.empty()) {
            int i = todo.back();
            todo.pop_back();
            nNodes = std::max<int64_t>(nNodes, i + 1);
            if (ln[i].nPrimitives == 0) {
                todo.push_back(i + 1);
                todo.push_back(ln[i].secondChildOffset);
            }
        }
    }
    static_assert(sizeof(LinearBVHNode) == sizeof(pb2_bvh_node), "pb2_bvh_node is the reference's LinearBVHNode");
    f->nodes.resize((size_t)nNodes);
    std::memcpy(f->nodes.data(), ln, (size_t)nNodes * sizeof(pb2_bvh_node));

    std::unordered_map<const Light *, int> lightNumber;
    for (size_t i = 0; i < scene.lights.size(); ++i) lightNumber[scene.lights[i].get()] = (int)i;
    f->lights.assign(scene.lights.size(), pb2_light{});
    std::unordered_map<const Material *, int> materialNumber;
    std::unordered_map<const TriangleMesh *, int> meshNumber;
    const size_t n = bvh->primitives.size();
    f->primType.resize(n);
    f->primIndex.resize(n);
    f->primMaterial.resize(n);
    f->primLight.assign(n, -1);
    f->bvhPrims.resize(n);
    bool constant = true;
    for (size_t i = 0; i < n; ++i) {   // primitives are numbered in BVHAccel::primitives order
        f->bvhPrims[i] = (int32_t)i;
        const GeometricPrimitive *gp = dynamic_cast<const GeometricPrimitive *>(bvh->primitives[i].get());
        if (!gp) {
            *err = "only GeometricPrimitives are bound by this overlay (object instances: see pbrt_v3_b200/csrc/host/accel.cpp)";
            return false;
        }
        // material
        int mid = -1;
        if (const Material *m = gp->material.get()) {
            auto it = materialNumber.find(m);
            if (it != materialNumber.end()) mid = it->second;
            else {
                pb2_material rec;
                std::memset(&rec, 0, sizeof(rec));
                if (const MatteMaterial *mm = dynamic_cast<const MatteMaterial *>(m)) {
                    rec.type = PB2_MAT_MATTE;
                    constantValue(mm->Kd, &constant).ToRGB(rec.kd);
                    rec.sigma = constantValue(mm->sigma, &constant);
                    if (mm->bumpMap) constant = false;
                } else if (const PlasticMaterial *pm = dynamic_cast<const PlasticMaterial *>(m)) {
                    rec.type = PB2_MAT_PLASTIC;
                    constantValue(pm->Kd, &constant).ToRGB(rec.kd);
                    constantValue(pm->Ks, &constant).ToRGB(rec.ks);
                    rec.roughness = constantValue(pm->roughness, &constant);
                    rec.remap_roughness = pm->remapRoughness ? 1 : 0;
                    if (pm->bumpMap) constant = false;
                } else {
                    *err = "this overlay binds matte and plastic (the other materials of the path: pbrt_v3_b200/csrc/host/render.cpp)";
                    return false;
                }
                mid = (int)f->materials.size();
                f->materials.push_back(rec);
                materialNumber[m] = mid;
            }
        }
        f->primMaterial[i] = mid;
        // area light
        if (const AreaLight *al = gp->areaLight.get()) {
            const DiffuseAreaLight *dl = dynamic_cast<const DiffuseAreaLight *>(al);
            auto it = lightNumber.find(al);
            if (!dl || it == lightNumber.end()) {
                *err = "area light that is not a DiffuseAreaLight of Scene::lights";
                return false;
            }
            pb2_light rec;
            std::memset(&rec, 0, sizeof(rec));
            rec.prim = (int32_t)i;
            dl->Lemit.ToRGB(rec.L);
            rec.two_sided = dl->twoSided ? 1 : 0;
            rec.area = dl->area;
            rec.type = PB2_LIGHT_AREA;
            f->lights[(size_t)it->second] = rec;
            f->primLight[i] = it->second;
        }
        // shape
        if (const Triangle *tri = dynamic_cast<const Triangle *>(gp->shape.get())) {
            const TriangleMesh *mesh = tri->mesh.get();
            auto it = meshNumber.find(mesh);
            int meshId;
            if (it != meshNumber.end()) meshId = it->second;
            else {
                meshId = (int)f->meshes.size();
                meshNumber[mesh] = meshId;
                pb2_mesh rec;
                std::memset(&rec, 0, sizeof(rec));
                rec.first_tri = -1;   // triangles of one mesh need not be contiguous in BVH order: unused by the library
                rec.n_tris = mesh->nTriangles;
                rec.first_vertex = (int32_t)(f->P.size() / 3);
                rec.n_vertices = mesh->nVertices;
                rec.has_n = mesh->n ? 1 : 0;
                rec.has_uv = mesh->uv ? 1 : 0;
                rec.has_s = mesh->s ? 1 : 0;
                rec.reverse_orientation = tri->reverseOrientation ? 1 : 0;
                rec.transform_swaps_handedness = tri->transformSwapsHandedness ? 1 : 0;
                for (int v = 0; v < mesh->nVertices; ++v) {
                    f->P.insert(f->P.end(), {mesh->p[v].x, mesh->p[v].y, mesh->p[v].z});
                    if (mesh->n) f->N.insert(f->N.end(), {mesh->n[v].x, mesh->n[v].y, mesh->n[v].z});
                    else f->N.insert(f->N.end(), {0.f, 0.f, 0.f});
                    if (mesh->uv) f->UV.insert(f->UV.end(), {mesh->uv[v].x, mesh->uv[v].y});
                    else f->UV.insert(f->UV.end(), {0.f, 0.f});
                    if (mesh->s) f->S.insert(f->S.end(), {mesh->s[v].x, mesh->s[v].y, mesh->s[v].z});
                    else f->S.insert(f->S.end(), {0.f, 0.f, 0.f});
                }
                f->meshes.push_back(rec);
                if (mesh->alphaMask || mesh->shadowAlphaMask) constant = false;
            }
            const int32_t base = f->meshes[(size_t)meshId].first_vertex;
            f->primType[i] = PB2_PRIM_TRIANGLE;
            f->primIndex[i] = (int32_t)f->triMesh.size();
            f->triMesh.push_back(meshId);
            for (int k = 0; k < 3; ++k) f->triIndex.push_back(base + tri->v[k]);
        } else if (const Sphere *sp = dynamic_cast<const Sphere *>(gp->shape.get())) {
            pb2_sphere rec;
            std::memset(&rec, 0, sizeof(rec));
            copyMatrix(sp->ObjectToWorld->GetMatrix(), rec.object_to_world);
            copyMatrix(sp->WorldToObject->GetMatrix(), rec.world_to_object);
            rec.radius = sp->radius;
            rec.z_min = sp->zMin;
            rec.z_max = sp->zMax;
            rec.theta_min = sp->thetaMin;
            rec.theta_max = sp->thetaMax;
            rec.phi_max = sp->phiMax;
            rec.reverse_orientation = sp->reverseOrientation ? 1 : 0;
            rec.transform_swaps_handedness = sp->transformSwapsHandedness ? 1 : 0;
            f->primType[i] = PB2_PRIM_SPHERE;
            f->primIndex[i] = (int32_t)f->spheres.size();
            f->spheres.push_back(rec);
        } else {
            *err = "shape outside the path's scope (triangle meshes and spheres)";
            return false;
        }
    }
    if (!constant) {
        *err = "a texture that varies over the surface / an alpha mask / a bump map: outside the path's scope";
        return false;
    }
    for (const pb2_light &l : f->lights)
        if (l.type != PB2_LIGHT_AREA || l.area == 0) {
            // (a record that was never filled: a light of the scene that is not attached to a primitive)
            *err = "this overlay binds DiffuseAreaLights (delta lights: pbrt_v3_b200/csrc/host/accel.cpp)";
            return false;
        }
    pb2_scene_desc &d = f->desc;
    std::memset(&d, 0, sizeof(d));
    d.n_vertices = (int64_t)(f->P.size() / 3);
    d.P = f->P.data();
    d.N = f->N.data();
    d.UV = f->UV.data();
    d.S = f->S.data();
    d.n_tris = (int64_t)f->triMesh.size();
    d.tri_index = f->triIndex.data();
    d.tri_mesh = f->triMesh.data();
    d.n_meshes = (int32_t)f->meshes.size();
    d.meshes = f->meshes.data();
    d.n_spheres = (int32_t)f->spheres.size();
    d.spheres = f->spheres.data();
    d.n_prims = (int64_t)n;
    d.prim_type = f->primType.data();
    d.prim_index = f->primIndex.data();
    d.prim_material = f->primMaterial.data();
    d.prim_light = f->primLight.data();
    d.n_nodes = nNodes;
    d.nodes = f->nodes.data();
    d.bvh_prims = f->bvhPrims.data();
    d.n_materials = (int32_t)f->materials.size();
    d.materials = f->materials.data();
    d.n_lights = (int32_t)f->lights.size();
    d.lights = f->lights.data();
    d.light_strategy = strategy == "uniform" ? PB2_LIGHTDIST_UNIFORM : strategy == "power" ? PB2_LIGHTDIST_POWER : PB2_LIGHTDIST_SPATIAL;
    d.spatial_max_voxels = 64;
    return true;
}

}  // namespace

void B200PathIntegrator::Render(const Scene &scene) {
    Flat flat;
    if (!FlattenForPb2(scene, lightStrategy, &flat, &error)) return;
    const PerspectiveCamera *pc = dynamic_cast<const PerspectiveCamera *>(camera.get());
    const HaltonSampler *hs = dynamic_cast<const HaltonSampler *>(sampler.get());
    Film *film = camera->film;
    const SobolSampler *ss = dynamic_cast<const SobolSampler *>(sampler.get());
    if (!pc || (!hs && !ss)) {
        error = "the path binds PerspectiveCamera and the two GlobalSamplers (HaltonSampler, SobolSampler)";
        return;
    }
    pb2_camera cam;
    std::memset(&cam, 0, sizeof(cam));
    copyMatrix(pc->CameraToWorld.startTransform->GetMatrix(), cam.camera_to_world);
    copyMatrix(pc->CameraToWorld.startTransform->GetInverseMatrix(), cam.world_to_camera);
    copyMatrix(pc->RasterToCamera.GetMatrix(), cam.raster_to_camera);
    cam.lens_radius = pc->lensRadius;
    cam.focal_distance = pc->focalDistance;
    cam.shutter_open = pc->shutterOpen;
    cam.shutter_close = pc->shutterClose;
    pb2_film_desc fd;
    std::memset(&fd, 0, sizeof(fd));
    fd.full_resolution[0] = film->fullResolution.x;
    fd.full_resolution[1] = film->fullResolution.y;
    fd.cropped_pixel_bounds[0] = film->croppedPixelBounds.pMin.x;
    fd.cropped_pixel_bounds[1] = film->croppedPixelBounds.pMin.y;
    fd.cropped_pixel_bounds[2] = film->croppedPixelBounds.pMax.x;
    fd.cropped_pixel_bounds[3] = film->croppedPixelBounds.pMax.y;
    fd.filter_radius[0] = film->filter->radius.x;
    fd.filter_radius[1] = film->filter->radius.y;
    fd.max_sample_luminance = film->maxSampleLuminance;
    fd.scale = film->scale;
    if (const GaussianFilter *g = dynamic_cast<const GaussianFilter *>(film->filter.get())) {
        fd.filter_type = PB2_FILTER_GAUSSIAN;
        fd.filter_param[0] = g->alpha;
    } else if (const MitchellFilter *m = dynamic_cast<const MitchellFilter *>(film->filter.get())) {
        fd.filter_type = PB2_FILTER_MITCHELL;
        fd.filter_param[0] = m->B;
        fd.filter_param[1] = m->C;
    } else if (const LanczosSincFilter *l = dynamic_cast<const LanczosSincFilter *>(film->filter.get())) {
        fd.filter_type = PB2_FILTER_SINC;
        fd.filter_param[0] = l->tau;
    } else if (dynamic_cast<const TriangleFilter *>(film->filter.get()))
        fd.filter_type = PB2_FILTER_TRIANGLE;
    else
        fd.filter_type = PB2_FILTER_BOX;
    pb2_path_params pp;
    std::memset(&pp, 0, sizeof(pp));
    pp.samples_per_pixel = (int32_t)sampler->samplesPerPixel;
    pp.sample_at_pixel_center = (hs && hs->sampleAtPixelCenter) ? 1 : 0;
    pp.sampler = ss ? PB2_SAMPLER_SOBOL : PB2_SAMPLER_HALTON;
    pp.max_depth = maxDepth;
    pp.rr_threshold = rrThreshold;
    pp.pixel_bounds[0] = pixelBounds.pMin.x;
    pp.pixel_bounds[1] = pixelBounds.pMin.y;
    pp.pixel_bounds[2] = pixelBounds.pMax.x;
    pp.pixel_bounds[3] = pixelBounds.pMax.y;
    pp.tile_count = 0;   // the library's own partition (all local GPUs / the communicator)

    if (pb2_device_count() == 0 && pb2_init(0) != PB2_OK) {   // no CPU fallback by design
        error = pb2_last_error();
        return;
    }
    pb2_scene *dev = nullptr;
    if (pb2_scene_create(&flat.desc, &dev) != PB2_OK) {
        error = pb2_last_error();
        return;
    }
    const Bounds2i cb = film->croppedPixelBounds;
    const int w = cb.pMax.x - cb.pMin.x;
    std::vector<float> rgbw((size_t)4 * cb.Area());
    if (pb2_render_path(dev, &cam, &fd, &pp, rgbw.data(), &stats) != PB2_OK)
        error = pb2_last_error();
    else {
        // hand the merged film to the reference's Film exactly as a tile would be merged (film.cpp:117-130)
        std::unique_ptr<FilmTile> tile = film->GetFilmTile(cb);
        for (Point2i p : tile->GetPixelBounds()) {
            if (p.x < cb.pMin.x || p.x >= cb.pMax.x || p.y < cb.pMin.y || p.y >= cb.pMax.y) continue;
            const float *v = &rgbw[4 * ((size_t)(p.y - cb.pMin.y) * w + (p.x - cb.pMin.x))];
            FilmTilePixel &px = tile->GetPixel(p);
            px.contribSum = Spectrum::FromRGB(v);
            px.filterWeightSum = v[3];
        }
        film->MergeFilmTile(std::move(tile));
        film->WriteImage();
    }
    pb2_scene_destroy(dev);
}

}  // namespace pbrt

// The reference Scene of handle `h` (built by ref_scene_create from a description, i.e. reference BVHAccel / Triangle /
// Material / Light objects) rendered by B200PathIntegrator.  out_rgb as ref_render.
extern "C" int ref_render_b200(void *h, const pb2_camera *cam, const pb2_film_desc *fd, const pb2_path_params *pp, float *out_rgb,
                               pb2_stats *stats, char *err, int err_len) {
    RefScene *rs = static_cast<RefScene *>(h);
    setThreads(1);
    RenderObjects ro = makeRenderObjects(*rs, cam, fd, pp);   // the reference's Film, PerspectiveCamera, HaltonSampler
    Bounds2i pb(Point2i(pp->pixel_bounds[0], pp->pixel_bounds[1]), Point2i(pp->pixel_bounds[2], pp->pixel_bounds[3]));
    B200PathIntegrator integrator(pp->max_depth, ro.camera, ro.sampler, pb, pp->rr_threshold, strategyName(rs->lightStrategy));
    integrator.Render(*rs->scene);
    if (!integrator.error.empty()) {
        if (err && err_len > 0) std::snprintf(err, (size_t)err_len, "%s", integrator.error.c_str());
        return 1;
    }
    {
        std::lock_guard<std::mutex> lock(g_imageMutex);
        if (out_rgb) std::memcpy(out_rgb, g_lastImage.data(), g_lastImage.size() * sizeof(float));
    }
    if (stats) *stats = integrator.stats;
    return 0;
}
