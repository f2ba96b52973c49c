// Developer tool (not part of the product, not used by tests): compiles the per-lane device headers
// with g++ and runs them on the CPU, to step through a parity difference without GPU time.
//   g++ -std=c++17 -O1 -g -ffp-contract=off -shared -fPIC -I../../include -I../../pbrt_v3_b200/csrc tools/emu/pb2_emu.cpp -o /tmp/libpb2_emu.so
#include <cstdint>
#include <cstring>
#include <vector>
struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
#include "device/pb2_path.cuh"
using namespace pb2;

struct EmuScene {
    DScene d;
    std::vector<float4> leaf;
};

extern "C" {
void *emu_scene_create(const pb2_scene_desc *desc) {
    EmuScene *s = new EmuScene;
    DScene &sc = s->d;
    std::memset(&sc, 0, sizeof(sc));
    sc.nNodes = desc->n_nodes; sc.nPrims = desc->n_prims; sc.nTris = desc->n_tris; sc.nLights = desc->n_lights;
    sc.nodes = reinterpret_cast<const float4 *>(desc->nodes);
    sc.P = desc->P; sc.N = desc->N; sc.UV = desc->UV; sc.S = desc->S;
    sc.triIndex = desc->tri_index; sc.triMesh = desc->tri_mesh; sc.meshes = desc->meshes; sc.spheres = desc->spheres;
    sc.primType = desc->prim_type; sc.primIndex = desc->prim_index; sc.primMaterial = desc->prim_material; sc.primLight = desc->prim_light;
    sc.materials = desc->materials; sc.lights = desc->lights;
    s->leaf.resize(3 * (size_t)desc->n_prims);
    for (int64_t j = 0; j < desc->n_prims; ++j) {
        int prim = desc->bvh_prims[j];
        float4 a, b, c;
        if (sc.primType[prim] == PB2_PRIM_SPHERE) {
            a = make_float4(0, 0, 0, bitsFloat((uint32_t)prim));
            b = make_float4(0, 0, 0, bitsFloat(LEAF_SPHERE));
            c = make_float4(0, 0, 0, bitsFloat((uint32_t)sc.primIndex[prim]));
        } else {
            int tri = sc.primIndex[prim];
            TriVerts t = triVerts(sc, tri);
            uint32_t flags = 0;
            V2 uv[3];
            triUVs(sc, tri, sc.meshes[sc.triMesh[tri]], uv);
            V3 dpdu, dpdv;
            if (!triPartials(t.p0, t.p1, t.p2, uv, &dpdu, &dpdv)) flags |= LEAF_DEGENERATE;
            a = make_float4(t.p0.x, t.p0.y, t.p0.z, bitsFloat((uint32_t)prim));
            b = make_float4(t.p1.x, t.p1.y, t.p1.z, bitsFloat(flags));
            c = make_float4(t.p2.x, t.p2.y, t.p2.z, 0.f);
        }
        s->leaf[3 * j] = a; s->leaf[3 * j + 1] = b; s->leaf[3 * j + 2] = c;
    }
    sc.leafPrims = s->leaf.data();
    return s;
}
int emu_intersect(void *h, const pb2_ray *rays, int64_t n, pb2_hit *hits) {
    const DScene &sc = static_cast<EmuScene *>(h)->d;
    for (int64_t i = 0; i < n; ++i) {
        DRay r;
        r.o = mk3(rays[i].o[0], rays[i].o[1], rays[i].o[2]);
        r.d = mk3(rays[i].d[0], rays[i].d[1], rays[i].d[2]);
        r.tMax = rays[i].t_max;
        DHit hit; hit.leaf = -1; hit.b0 = hit.b1 = hit.b2 = 0;
        float tMax = r.tMax;
        bool found = traverse<false>(sc, r, &tMax, &hit, nullptr);
        pb2_hit out; std::memset(&out, 0, sizeof(out));
        out.prim = -1; out.t = tMax;
        if (found) {
            DInteraction it = hitInteraction<true>(sc, hit, r, tMax);
            out.prim = it.prim;
            out.b[0] = hit.b0; out.b[1] = hit.b1; out.b[2] = hit.b2;
            out.p[0] = it.p.x; out.p[1] = it.p.y; out.p[2] = it.p.z;
            out.p_error[0] = it.pError.x; out.p_error[1] = it.pError.y; out.p_error[2] = it.pError.z;
            out.n[0] = it.n.x; out.n[1] = it.n.y; out.n[2] = it.n.z;
            out.ns[0] = it.ns.x; out.ns[1] = it.ns.y; out.ns[2] = it.ns.z;
            out.dpdu[0] = it.dpdus.x; out.dpdu[1] = it.dpdus.y; out.dpdu[2] = it.dpdus.z;
            out.uv[0] = it.uv.x; out.uv[1] = it.uv.y;
        }
        hits[i] = out;
    }
    return 0;
}
}
