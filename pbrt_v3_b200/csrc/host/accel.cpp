// Host BVH construction (the traversal lives on the GPU: ../device/traverse.cuh).
//
// Produces the reference's 32-byte depth-first LinearBVHNode array (src/accelerators/bvh.cpp:95-104,
// 640-658) with the same split decisions as BVHAccel::recursiveBuild (bvh.cpp:236-402): 12-bucket
// SAH, leaf when cheaper and <= maxPrimsInNode, middle / equal-count alternatives.  The tree is
// built directly into a flat pool of build records instead of arena-allocated pointer nodes, and
// the primitive order is produced as an index permutation.
#include <atomic>
#include <chrono>
#include <functional>
#include <future>
#include <thread>
#include <unordered_map>

#include "scene.h"

namespace pbrt {

namespace {
struct PrimInfo {
    size_t number;
    Bounds3f bounds;
    Point3f centroid;
};
struct BuildNode {
    Bounds3f bounds;
    int child[2] = {-1, -1};
    int splitAxis = 0, firstPrimOffset = 0, nPrimitives = 0;
};
struct Builder {
    std::vector<PrimInfo> info;
    std::vector<BuildNode> pool;
    std::vector<int32_t> ordered;
    int maxPrimsInNode;
    BVHAccel::SplitMethod method;

    // The build runs on several threads: sub-ranges of `info` are disjoint, build records come from a
    // pre-sized pool through an atomic counter, and a subtree over n primitives owns a fixed range of
    // n slots of the ordered list (orderedStart), so the result does not depend on the schedule.
    std::atomic<int> nextNode{0};

    int makeLeaf(int node, int start, int end, const Bounds3f &bounds, int orderedStart) {
        pool[node].firstPrimOffset = orderedStart;
        for (int i = start; i < end; ++i) ordered[orderedStart + (i - start)] = (int32_t)info[i].number;
        pool[node].nPrimitives = end - start;
        pool[node].bounds = bounds;
        return node;
    }

    int build(int start, int end, int orderedStart, int spawnLevels) {
        int node = nextNode.fetch_add(1);
        Bounds3f bounds;
        for (int i = start; i < end; ++i) bounds = Union(bounds, info[i].bounds);
        int nPrims = end - start;
        if (nPrims == 1) return makeLeaf(node, start, end, bounds, orderedStart);
        Bounds3f cb;
        for (int i = start; i < end; ++i) cb = Union(cb, info[i].centroid);
        int dim = cb.MaximumExtent();
        if (cb.pMax[dim] == cb.pMin[dim]) return makeLeaf(node, start, end, bounds, orderedStart);

        int mid = (start + end) / 2;
        auto equalCounts = [&]() {
            mid = (start + end) / 2;
            std::nth_element(&info[start], &info[mid], &info[end - 1] + 1,
                             [dim](const PrimInfo &a, const PrimInfo &b) { return a.centroid[dim] < b.centroid[dim]; });
        };
        bool done = false;
        if (method == BVHAccel::SplitMethod::Middle) {
            Float pmid = (cb.pMin[dim] + cb.pMax[dim]) / 2;
            PrimInfo *m = std::partition(&info[start], &info[end - 1] + 1,
                                         [dim, pmid](const PrimInfo &pi) { return pi.centroid[dim] < pmid; });
            mid = (int)(m - &info[0]);
            done = (mid != start && mid != end);
            if (!done) equalCounts();  // bvh.cpp:286-290 falls through to EqualCounts
        } else if (method == BVHAccel::SplitMethod::EqualCounts) {
            equalCounts();
        } else {
            if (nPrims <= 2) {
                equalCounts();
            } else {
                constexpr int nBuckets = 12;
                struct Bucket { int count = 0; Bounds3f bounds; } buckets[nBuckets];
                auto bucketOf = [&](const PrimInfo &pi) {
                    int b = nBuckets * cb.Offset(pi.centroid)[dim];
                    if (b == nBuckets) b = nBuckets - 1;
                    return b;
                };
                for (int i = start; i < end; ++i) {
                    int b = bucketOf(info[i]);
                    buckets[b].count++;
                    buckets[b].bounds = Union(buckets[b].bounds, info[i].bounds);
                }
                Float cost[nBuckets - 1];
                for (int i = 0; i < nBuckets - 1; ++i) {
                    Bounds3f b0, b1;
                    int c0 = 0, c1 = 0;
                    for (int j = 0; j <= i; ++j) { b0 = Union(b0, buckets[j].bounds); c0 += buckets[j].count; }
                    for (int j = i + 1; j < nBuckets; ++j) { b1 = Union(b1, buckets[j].bounds); c1 += buckets[j].count; }
                    cost[i] = 1 + (c0 * b0.SurfaceArea() + c1 * b1.SurfaceArea()) / bounds.SurfaceArea();
                }
                Float minCost = cost[0];
                int minBucket = 0;
                for (int i = 1; i < nBuckets - 1; ++i)
                    if (cost[i] < minCost) { minCost = cost[i]; minBucket = i; }
                Float leafCost = nPrims;
                if (nPrims > maxPrimsInNode || minCost < leafCost) {
                    PrimInfo *m = std::partition(&info[start], &info[end - 1] + 1,
                                                 [&](const PrimInfo &pi) { return bucketOf(pi) <= minBucket; });
                    mid = (int)(m - &info[0]);
                } else
                    return makeLeaf(node, start, end, bounds, orderedStart);
            }
        }
        // The reference passes both recursive calls as arguments of InitInterior (bvh.cpp:394-398);
        // its compiler (gcc, x86-64) evaluates them right to left, so the second child's primitives
        // are appended to the ordered list first.  Keep that order: it fixes primitivesOffset values.
        // -> the second child's primitives come first in the ordered list: its range starts at
        // orderedStart, the first child's right after it.
        int c0, c1;
        if (spawnLevels > 0 && nPrims > 65536) {
            std::future<int> second = std::async(std::launch::async, [=] { return build(mid, end, orderedStart, spawnLevels - 1); });
            c0 = build(start, mid, orderedStart + (end - mid), spawnLevels - 1);
            c1 = second.get();
        } else {
            c1 = build(mid, end, orderedStart, 0);
            c0 = build(start, mid, orderedStart + (end - mid), 0);
        }
        pool[node].child[0] = c0;
        pool[node].child[1] = c1;
        pool[node].bounds = Union(pool[c0].bounds, pool[c1].bounds);
        pool[node].splitAxis = dim;
        pool[node].nPrimitives = 0;
        return node;
    }

    // ---- HLBVH (bvh.cpp:404-638): Morton-ordered treelets below, a SAH tree over the treelet roots above.
    struct MortonPrim { int32_t primitiveIndex; uint32_t mortonCode; };
    static uint32_t leftShift3(uint32_t x) {   // bvh.cpp:106-130
        if (x == (1 << 10)) --x;
        x = (x | (x << 16)) & 0x30000ff;
        x = (x | (x << 8)) & 0x300f00f;
        x = (x | (x << 4)) & 0x30c30c3;
        x = (x | (x << 2)) & 0x9249249;
        return x;
    }
    // emitLBVH (bvh.cpp:472-539).  `first` is the position of mp[0] in the sorted list: the reference hands out the
    // ordered-primitive slots through an atomic counter as leaves are created; run by one thread that is the sorted
    // order itself, which is what every schedule produces here.
    int emitLBVH(const MortonPrim *mp, int first, int nPrimitives, int bitIndex) {
        if (bitIndex == -1 || nPrimitives < maxPrimsInNode) {
            int node = nextNode.fetch_add(1);
            Bounds3f bounds;
            for (int i = 0; i < nPrimitives; ++i) {
                ordered[first + i] = mp[i].primitiveIndex;
                bounds = Union(bounds, info[mp[i].primitiveIndex].bounds);
            }
            pool[node].firstPrimOffset = first;
            pool[node].nPrimitives = nPrimitives;
            pool[node].bounds = bounds;
            return node;
        }
        uint32_t mask = 1u << bitIndex;
        if ((mp[0].mortonCode & mask) == (mp[nPrimitives - 1].mortonCode & mask)) return emitLBVH(mp, first, nPrimitives, bitIndex - 1);
        int searchStart = 0, searchEnd = nPrimitives - 1;
        while (searchStart + 1 != searchEnd) {
            int mid = (searchStart + searchEnd) / 2;
            if ((mp[searchStart].mortonCode & mask) == (mp[mid].mortonCode & mask)) searchStart = mid;
            else searchEnd = mid;
        }
        int splitOffset = searchEnd;
        int node = nextNode.fetch_add(1);
        int c0 = emitLBVH(mp, first, splitOffset, bitIndex - 1);
        int c1 = emitLBVH(mp + splitOffset, first + splitOffset, nPrimitives - splitOffset, bitIndex - 1);
        pool[node].child[0] = c0;
        pool[node].child[1] = c1;
        pool[node].bounds = Union(pool[c0].bounds, pool[c1].bounds);
        pool[node].splitAxis = bitIndex % 3;
        pool[node].nPrimitives = 0;
        return node;
    }
    // buildUpperSAH (bvh.cpp:541-638)
    int buildUpperSAH(std::vector<int> &roots, int start, int end) {
        int nNodes = end - start;
        if (nNodes == 1) return roots[start];
        int node = nextNode.fetch_add(1);
        Bounds3f bounds;
        for (int i = start; i < end; ++i) bounds = Union(bounds, pool[roots[i]].bounds);
        Bounds3f centroidBounds;
        for (int i = start; i < end; ++i) {
            Point3f centroid = (pool[roots[i]].bounds.pMin + pool[roots[i]].bounds.pMax) * 0.5f;
            centroidBounds = Union(centroidBounds, centroid);
        }
        int dim = centroidBounds.MaximumExtent();
        constexpr int nBuckets = 12;
        struct Bucket { int count = 0; Bounds3f bounds; } buckets[nBuckets];
        const Float cmin = centroidBounds.pMin[dim], cmax = centroidBounds.pMax[dim];
        auto bucketOf = [&](int r) {
            Float centroid = (pool[r].bounds.pMin[dim] + pool[r].bounds.pMax[dim]) * 0.5f;
            int b = nBuckets * ((centroid - cmin) / (cmax - cmin));
            if (b == nBuckets) b = nBuckets - 1;
            return b;
        };
        for (int i = start; i < end; ++i) {
            int b = bucketOf(roots[i]);
            buckets[b].count++;
            buckets[b].bounds = Union(buckets[b].bounds, pool[roots[i]].bounds);
        }
        Float cost[nBuckets - 1];
        for (int i = 0; i < nBuckets - 1; ++i) {
            Bounds3f b0, b1;
            int count0 = 0, count1 = 0;
            for (int j = 0; j <= i; ++j) { b0 = Union(b0, buckets[j].bounds); count0 += buckets[j].count; }
            for (int j = i + 1; j < nBuckets; ++j) { b1 = Union(b1, buckets[j].bounds); count1 += buckets[j].count; }
            cost[i] = .125f + (count0 * b0.SurfaceArea() + count1 * b1.SurfaceArea()) / bounds.SurfaceArea();
        }
        Float minCost = cost[0];
        int minCostSplitBucket = 0;
        for (int i = 1; i < nBuckets - 1; ++i)
            if (cost[i] < minCost) { minCost = cost[i]; minCostSplitBucket = i; }
        int *pmid = std::partition(&roots[start], &roots[end - 1] + 1, [&](int r) { return bucketOf(r) <= minCostSplitBucket; });
        int mid = (int)(pmid - &roots[0]);
        int c0 = buildUpperSAH(roots, start, mid);
        int c1 = buildUpperSAH(roots, mid, end);
        pool[node].child[0] = c0;
        pool[node].child[1] = c1;
        pool[node].bounds = Union(pool[c0].bounds, pool[c1].bounds);
        pool[node].splitAxis = dim;
        pool[node].nPrimitives = 0;
        return node;
    }
    // The whole HLBVH build on the GPU (pb2_hlbvh_build, include/pb2.h): Morton codes, sort, treelets, the SAH tree over
    // their roots and the depth-first layout; the finished LinearBVHNode array and the primitive order come back.
    bool buildHLBVHAllOnDevice(std::vector<pb2_bvh_node> *nodes, double *deviceMs) {
        const int n = (int)info.size();
        std::vector<float> primBounds((size_t)n * 6);
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < 3; ++k) {
                primBounds[6 * (size_t)i + k] = info[i].bounds.pMin[k];
                primBounds[6 * (size_t)i + 3 + k] = info[i].bounds.pMax[k];
            }
        if (!EnsureDevice()) return false;
        nodes->resize((size_t)2 * n);
        int64_t nNodes = 0;
        if (pb2_hlbvh_build(primBounds.data(), n, maxPrimsInNode, nodes->data(), &nNodes, ordered.data(), deviceMs) != PB2_OK) {
            Error("Device BVH build failed: %s", pb2_last_error());
            nodes->clear();
            return false;
        }
        nodes->resize((size_t)nNodes);
        return true;
    }
    // Only the lower half on the GPU (pb2_hlbvh_treelets; PB2_DEVICE_BVH_UPPER=host): Morton codes, sort and treelets come
    // back as this builder's own records, the small SAH tree over the treelet roots and the flatten run here.
    int buildHLBVHOnDevice(double *deviceMs) {
        static_assert(sizeof(BuildNode) == sizeof(pb2_build_node), "BuildNode is the ABI's pb2_build_node");
        const int n = (int)info.size();
        std::vector<float> primBounds((size_t)n * 6);
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < 3; ++k) {
                primBounds[6 * (size_t)i + k] = info[i].bounds.pMin[k];
                primBounds[6 * (size_t)i + 3 + k] = info[i].bounds.pMax[k];
            }
        std::vector<int32_t> roots32(4096);
        int32_t nTreelets = 0;
        if (!EnsureDevice()) return -1;
        pool.resize((size_t)2 * n + 4096);
        if (pb2_hlbvh_treelets(primBounds.data(), n, maxPrimsInNode, reinterpret_cast<pb2_build_node *>(pool.data()), ordered.data(),
                               roots32.data(), &nTreelets, deviceMs) != PB2_OK) {
            Error("Device BVH build failed: %s", pb2_last_error());
            return -1;
        }
        nextNode.store(2 * n);
        std::vector<int> roots(roots32.begin(), roots32.begin() + nTreelets);
        return buildUpperSAH(roots, 0, (int)roots.size());
    }
    int buildHLBVH() {
        Bounds3f bounds;
        for (const PrimInfo &pi : info) bounds = Union(bounds, pi.centroid);
        const int n = (int)info.size();
        std::vector<MortonPrim> mortonPrims(n);
        const unsigned nThreads = std::max(1u, std::thread::hardware_concurrency());
        auto parallelFor = [&](int count, const std::function<void(int)> &fn) {
            if (count < 4096 || nThreads == 1) {
                for (int i = 0; i < count; ++i) fn(i);
                return;
            }
            std::atomic<int> next{0};
            std::vector<std::thread> workers;
            for (unsigned t = 0; t < nThreads; ++t)
                workers.emplace_back([&] {
                    for (;;) {
                        int b = next.fetch_add(1024);
                        if (b >= count) return;
                        for (int i = b; i < std::min(count, b + 1024); ++i) fn(i);
                    }
                });
            for (auto &w : workers) w.join();
        };
        parallelFor(n, [&](int i) {
            constexpr int mortonScale = 1 << 10;
            mortonPrims[i].primitiveIndex = (int32_t)info[i].number;
            Vector3f o = bounds.Offset(info[i].centroid) * mortonScale;
            mortonPrims[i].mortonCode = (leftShift3((uint32_t)o.z) << 2) | (leftShift3((uint32_t)o.y) << 1) | leftShift3((uint32_t)o.x);
        });
        // RadixSort (bvh.cpp:139-180) is a least-significant-digit sort over all 30 bits: a stable sort by the code
        std::stable_sort(mortonPrims.begin(), mortonPrims.end(), [](const MortonPrim &a, const MortonPrim &b) { return a.mortonCode < b.mortonCode; });
        struct Treelet { int start, nPrimitives, root; };
        std::vector<Treelet> treelets;
        for (int start = 0, end = 1; end <= n; ++end) {
            const uint32_t mask = 0x3ffc0000;
            if (end == n || ((mortonPrims[start].mortonCode & mask) != (mortonPrims[end].mortonCode & mask))) {
                treelets.push_back({start, end - start, -1});
                start = end;
            }
        }
        parallelFor((int)treelets.size(), [&](int i) {
            Treelet &tr = treelets[i];
            tr.root = emitLBVH(&mortonPrims[tr.start], tr.start, tr.nPrimitives, 29 - 12);
        });
        std::vector<int> roots;
        roots.reserve(treelets.size());
        for (const Treelet &tr : treelets) roots.push_back(tr.root);
        return buildUpperSAH(roots, 0, (int)roots.size());
    }

    int flatten(int bn, std::vector<pb2_bvh_node> &out) {
        int my = (int)out.size();
        out.push_back(pb2_bvh_node());
        const BuildNode &n = pool[bn];
        pb2_bvh_node rec;
        std::memset(&rec, 0, sizeof(rec));
        for (int k = 0; k < 3; ++k) { rec.bmin[k] = n.bounds.pMin[k]; rec.bmax[k] = n.bounds.pMax[k]; }
        if (n.nPrimitives > 0) {
            rec.offset = n.firstPrimOffset;
            rec.n_prims = (uint16_t)n.nPrimitives;
            out[my] = rec;
        } else {
            rec.axis = (uint8_t)n.splitAxis;
            rec.n_prims = 0;
            out[my] = rec;
            flatten(n.child[0], out);
            int second = flatten(n.child[1], out);
            out[my].offset = second;
        }
        return my;
    }
};
}  // namespace

BVHAccel::BVHAccel(std::vector<std::shared_ptr<Primitive>> p, int maxPrims, SplitMethod sm, bool deviceBuild)
    : sceneOrderPrims(std::move(p)), maxPrimsInNode(std::min(255, maxPrims)), splitMethod(sm), deviceBuild(deviceBuild) {
    if (sceneOrderPrims.empty()) return;
    Builder b;
    b.maxPrimsInNode = maxPrimsInNode;
    b.method = splitMethod;
    auto clk = [] { return std::chrono::steady_clock::now(); };
    auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    const auto t0 = clk();
    b.info.resize(sceneOrderPrims.size());
    for (size_t i = 0; i < sceneOrderPrims.size(); ++i) {
        Bounds3f wb = sceneOrderPrims[i]->WorldBound();
        b.info[i] = PrimInfo{i, wb, .5f * wb.pMin + .5f * wb.pMax};
    }
    b.pool.resize(2 * sceneOrderPrims.size());
    b.ordered.resize(sceneOrderPrims.size());
    int spawnLevels = 0;
    for (unsigned t = std::max(1u, std::thread::hardware_concurrency()); t > 1; t >>= 1) ++spawnLevels;
    const auto t1 = clk();
    int root = -1;
    double deviceMs = 0;
    const char *upperEnv = std::getenv("PB2_DEVICE_BVH_UPPER");
    const bool upperOnHost = upperEnv && std::string(upperEnv) == "host";
    if (splitMethod == SplitMethod::HLBVH && deviceBuild && !upperOnHost) {
        // everything on the device: nothing left to flatten here
        if (!b.buildHLBVHAllOnDevice(&nodes, &deviceMs)) {
            sceneOrderPrims.clear();
            return;
        }
        lastBuildDeviceMs = deviceMs;
        orderedPrimNumbers = std::move(b.ordered);
        primitives.reserve(orderedPrimNumbers.size());
        for (int32_t n : orderedPrimNumbers) primitives.push_back(sceneOrderPrims[n]);
        return;
    }
    if (splitMethod == SplitMethod::HLBVH && deviceBuild) root = b.buildHLBVHOnDevice(&deviceMs);
    else if (splitMethod == SplitMethod::HLBVH) root = b.buildHLBVH();
    else root = b.build(0, (int)sceneOrderPrims.size(), 0, spawnLevels + 2);
    if (root < 0) {   // the device build reported an error: leave an empty accelerator, like a scene without primitives
        sceneOrderPrims.clear();
        return;
    }
    lastBuildDeviceMs = deviceMs;
    const auto t2 = clk();
    b.pool.resize((size_t)b.nextNode.load());
    nodes.reserve(b.pool.size());
    b.flatten(root, nodes);
    const auto t3 = clk();
    orderedPrimNumbers = std::move(b.ordered);
    primitives.reserve(orderedPrimNumbers.size());
    for (int32_t n : orderedPrimNumbers) primitives.push_back(sceneOrderPrims[n]);
    if (std::getenv("PB2_VERBOSE") && sceneOrderPrims.size() > 100000)
        std::fprintf(stderr, "pb2: BVHAccel: bounds %.2f s, recursive build %.2f s, flatten %.2f s, ordered primitives %.2f s\n",
                     secs(t0, t1), secs(t1, t2), secs(t2, t3), secs(t3, clk()));
}

BVHAccel::~BVHAccel() {}

Bounds3f BVHAccel::WorldBound() const {
    if (nodes.empty()) return Bounds3f();
    Bounds3f b;
    b.pMin = Point3f(nodes[0].bmin[0], nodes[0].bmin[1], nodes[0].bmin[2]);
    b.pMax = Point3f(nodes[0].bmax[0], nodes[0].bmax[1], nodes[0].bmax[2]);
    return b;
}

std::shared_ptr<BVHAccel> CreateBVHAccelerator(std::vector<std::shared_ptr<Primitive>> prims, const ParamSet &ps) {
    std::string name = ps.FindOneString("splitmethod", "sah");
    BVHAccel::SplitMethod sm;
    if (name == "sah") sm = BVHAccel::SplitMethod::SAH;
    else if (name == "hlbvh") sm = BVHAccel::SplitMethod::HLBVH;
    else if (name == "middle") sm = BVHAccel::SplitMethod::Middle;
    else if (name == "equal") sm = BVHAccel::SplitMethod::EqualCounts;
    else {
        Warning("BVH split method \"%s\" unknown.  Using \"sah\".", name.c_str());
        sm = BVHAccel::SplitMethod::SAH;
    }
    int maxPrimsInNode = ps.FindOneInt("maxnodeprims", 4);
    // Not a parameter of the reference: "bool devicebuild" (or PB2_DEVICE_BVH=1) asks for the HLBVH treelets to be
    // built by the CUDA library instead of the host threads; the tree is the same either way.
    const char *env = std::getenv("PB2_DEVICE_BVH");
    bool deviceBuild = ps.FindOneBool("devicebuild", env && env[0] == '1');
    if (deviceBuild && sm != BVHAccel::SplitMethod::HLBVH) {
        Warning("\"devicebuild\" applies to splitmethod \"hlbvh\" only; building on the host");
        deviceBuild = false;
    }
    return std::make_shared<BVHAccel>(std::move(prims), maxPrimsInNode, sm, deviceBuild);
}

const AreaLight *Aggregate::GetAreaLight() const {
    Error("Aggregate::GetAreaLight() method called; should have gone to GeometricPrimitive");
    return nullptr;
}
const Material *Aggregate::GetMaterial() const {
    Error("Aggregate::GetMaterial() method called; should have gone to GeometricPrimitive");
    return nullptr;
}

// ---------------------------------------------------------------- flatten: object graph -> pb2_scene_desc
static void copyMatrix(const Matrix4x4 &m, float out[16]) { std::memcpy(out, m.m, 16 * sizeof(float)); }

std::unique_ptr<FlatScene> FlattenScene(const BVHAccel &bvh, const std::vector<std::shared_ptr<Light>> &lights,
                                        const std::string &lightStrategy) {
    std::unique_ptr<FlatScene> fs(new FlatScene());
    std::unordered_map<const TriangleMesh *, int> meshIds;
    std::unordered_map<const Material *, int> materialIds;
    std::unordered_map<const AreaLight *, int> lightIds;
    // Scene::lights order is the order of creation in the scene file (api.cpp:1413-1416)
    fs->lights.resize(lights.size());
    std::vector<char> lightSeen(lights.size(), 0);
    // image textures in order of first use; 0 = none, else 1 + index (pb2_material::tex, pb2_mesh::alpha_tex)
    std::unordered_map<const ImageTexture *, int> textureIds;
    std::function<int32_t(const std::shared_ptr<ImageTexture> &)> textureId = [&](const std::shared_ptr<ImageTexture> &t) -> int32_t {
        if (!t) return 0;
        auto it = textureIds.find(t.get());
        if (it != textureIds.end()) return it->second;
        for (const auto &c : t->child) textureId(c);   // children before their parent (pb2_texture::child)
        fs->textureObjects.push_back(t);
        const int id = (int)fs->textureObjects.size();
        textureIds[t.get()] = id;
        return id;
    };
    for (size_t i = 0; i < lights.size(); ++i) {
        if (const AreaLight *al = dynamic_cast<const AreaLight *>(lights[i].get())) {
            lightIds[al] = (int)i;
            continue;
        }
        // delta lights: no primitive, geometry in the parallel record
        if (fs->deltaLights.empty()) {
            fs->deltaLights.resize(lights.size());
            std::memset(fs->deltaLights.data(), 0, fs->deltaLights.size() * sizeof(pb2_delta_light));
        }
        pb2_light rec;
        std::memset(&rec, 0, sizeof(rec));
        rec.prim = -1;
        pb2_delta_light &dl = fs->deltaLights[i];
        if (const PointLight *pl = dynamic_cast<const PointLight *>(lights[i].get())) {
            rec.type = PB2_LIGHT_POINT;
            for (int c = 0; c < 3; ++c) rec.L[c] = pl->I.c[c];
            dl.p[0] = pl->pLight.x; dl.p[1] = pl->pLight.y; dl.p[2] = pl->pLight.z;
        } else if (const SpotLight *sl = dynamic_cast<const SpotLight *>(lights[i].get())) {
            rec.type = PB2_LIGHT_SPOT;
            for (int c = 0; c < 3; ++c) rec.L[c] = sl->I.c[c];
            dl.p[0] = sl->pLight.x; dl.p[1] = sl->pLight.y; dl.p[2] = sl->pLight.z;
            dl.total_width_deg = sl->totalWidth;
            dl.falloff_start_deg = sl->falloffStart;
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) dl.world_to_light[3 * r + c] = sl->WorldToLight.GetMatrix().m[r][c];
        } else if (const DistantLight *dd = dynamic_cast<const DistantLight *>(lights[i].get())) {
            rec.type = PB2_LIGHT_DISTANT;
            for (int c = 0; c < 3; ++c) rec.L[c] = dd->L.c[c];
            dl.p[0] = dd->wWorld.x; dl.p[1] = dd->wWorld.y; dl.p[2] = dd->wWorld.z;
            // DistantLight::Preprocess (distant.h:55-57): Bounds3::BoundingSphere of Scene::WorldBound (geometry.h:769-772)
            Bounds3f wb = bvh.WorldBound();
            Point3f center = (wb.pMin + wb.pMax) / 2;
            bool inside = center.x >= wb.pMin.x && center.x <= wb.pMax.x && center.y >= wb.pMin.y && center.y <= wb.pMax.y &&
                          center.z >= wb.pMin.z && center.z <= wb.pMax.z;
            dl.world_radius = inside ? Distance(center, wb.pMax) : 0;
        } else if (const InfiniteAreaLight *il = dynamic_cast<const InfiniteAreaLight *>(lights[i].get())) {
            rec.type = PB2_LIGHT_INFINITE;
            for (int c = 0; c < 3; ++c) rec.L[c] = il->L.c[c];
            dl.env_tex = textureId(il->envMap);
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) {
                    dl.world_to_light[3 * r + c] = il->WorldToLight.GetMatrix().m[r][c];
                    dl.light_to_world[3 * r + c] = il->LightToWorld.GetMatrix().m[r][c];
                }
            // InfiniteAreaLight::Preprocess (infinite.h:61-63): the bounding sphere of Scene::WorldBound, as for DistantLight
            Bounds3f wb = bvh.WorldBound();
            Point3f center = (wb.pMin + wb.pMax) / 2;
            bool inside = center.x >= wb.pMin.x && center.x <= wb.pMax.x && center.y >= wb.pMin.y && center.y <= wb.pMax.y &&
                          center.z >= wb.pMin.z && center.z <= wb.pMax.z;
            dl.world_radius = inside ? Distance(center, wb.pMax) : 0;
        } else {
            Error("Light type outside the GPU path's scope (diffuse area, point, spot, distant, infinite; SURVEY.md §2 rows 14-15)");
            return nullptr;
        }
        fs->lights[i] = rec;
        lightSeen[i] = 1;
    }
    bool anyN = false, anyUV = false, anyS = false;

    // Primitive numbering: the scene-level primitives in scene order (GeometricPrimitives and
    // TransformedPrimitives), then the GeometricPrimitives of every distinct instanced object.
    std::vector<const Primitive *> &objs = fs->primObjects;
    for (const auto &p : bvh.sceneOrderPrims) objs.push_back(p.get());
    const size_t nTop = objs.size();
    struct ObjectInfo { int bvh; int32_t loneNumber; };
    std::unordered_map<const Primitive *, ObjectInfo> objects;
    std::vector<const BVHAccel *> objectBvhs;
    std::vector<size_t> objectBase;
    for (size_t i = 0; i < nTop; ++i) {
        const TransformedPrimitive *tp = dynamic_cast<const TransformedPrimitive *>(objs[i]);
        if (!tp) continue;
        const Primitive *inner = tp->primitive.get();
        if (objects.count(inner)) continue;
        if (const BVHAccel *ob = dynamic_cast<const BVHAccel *>(inner)) {
            objects[inner] = ObjectInfo{(int)objectBvhs.size() + 1, -1};
            objectBvhs.push_back(ob);
            objectBase.push_back(objs.size());
            for (const auto &p : ob->sceneOrderPrims) objs.push_back(p.get());
        } else if (dynamic_cast<const GeometricPrimitive *>(inner)) {
            objects[inner] = ObjectInfo{-1, (int32_t)objs.size()};
            objs.push_back(inner);
        } else {
            Error("An instanced object must be a GeometricPrimitive or a BVHAccel over GeometricPrimitives");
            return nullptr;
        }
    }
    size_t nPrims = objs.size();
    fs->primType.resize(nPrims);
    fs->primIndex.resize(nPrims);
    fs->primMaterial.resize(nPrims);
    fs->primLight.resize(nPrims);
    for (size_t i = 0; i < nPrims; ++i) {
        if (const TransformedPrimitive *tp = dynamic_cast<const TransformedPrimitive *>(objs[i])) {
            if (i >= nTop) {
                Error("Instances inside an instanced object are not allowed (api.cpp:1554-1557)");
                return nullptr;
            }
            pb2_instance inst;
            std::memset(&inst, 0, sizeof(inst));
            copyMatrix(tp->PrimitiveToWorld.GetMatrix(), inst.instance_to_world);
            copyMatrix(tp->PrimitiveToWorld.GetInverseMatrix(), inst.world_to_instance);
            const ObjectInfo &oi = objects[tp->primitive.get()];
            inst.bvh = oi.bvh;
            inst.lone_prim = oi.loneNumber;   // primitive number for now; becomes its bvh_prims position below
            fs->primType[i] = PB2_PRIM_INSTANCE;
            fs->primIndex[i] = (int32_t)fs->instances.size();
            fs->primMaterial[i] = -1;
            fs->primLight[i] = -1;
            fs->instances.push_back(inst);
            continue;
        }
        const GeometricPrimitive *gp = dynamic_cast<const GeometricPrimitive *>(objs[i]);
        if (!gp) {
            Error("Only GeometricPrimitives and TransformedPrimitives are supported below the scene BVH");
            return nullptr;
        }
        const bool insideObject = i >= nTop;
        // material
        int mid = -1;
        if (gp->material) {
            auto it = materialIds.find(gp->material.get());
            if (it == materialIds.end()) {
                mid = (int)fs->materials.size();
                pb2_material rec = gp->material->Record();
                for (int k = 0; k < PB2_TEX_SLOTS; ++k) rec.tex[k] = textureId(gp->material->tex[k]);
                fs->materials.push_back(rec);
                materialIds[gp->material.get()] = mid;
            } else
                mid = it->second;
        }
        fs->primMaterial[i] = mid;
        // light
        int lid = -1;
        if (gp->areaLight && insideObject) {
            // The reference warns "Area lights not supported with object instancing" (api.cpp:1408-1409) and
            // leaves the light out of Scene::lights; the primitive would still emit when seen directly.  That
            // unsampled emission is not carried to the device.
            Warning("Emission of an area light inside an instanced object is ignored on the GPU path");
        } else if (gp->areaLight && lights.empty()) {
            // an aggregate asked to intersect on its own (BVHAccel::Intersect before / without Render) is flattened
            // without lights: geometry only, emission is not part of that query
        } else if (gp->areaLight) {
            auto it = lightIds.find(gp->areaLight.get());
            if (it == lightIds.end()) {
                Error("Primitive's area light is not in the scene's light list");
                return nullptr;
            }
            lid = it->second;
            const DiffuseAreaLight *dl = dynamic_cast<const DiffuseAreaLight *>(gp->areaLight.get());
            pb2_light rec;
            std::memset(&rec, 0, sizeof(rec));
            rec.prim = (int32_t)i;
            rec.L[0] = dl->Lemit.c[0]; rec.L[1] = dl->Lemit.c[1]; rec.L[2] = dl->Lemit.c[2];
            rec.two_sided = dl->twoSided ? 1 : 0;
            rec.area = dl->area;
            fs->lights[lid] = rec;
            lightSeen[lid] = 1;
        }
        fs->primLight[i] = lid;
        // shape
        if (const Triangle *tri = dynamic_cast<const Triangle *>(gp->shape.get())) {
            const TriangleMesh *mesh = tri->mesh.get();
            auto it = meshIds.find(mesh);
            int meshId;
            if (it == meshIds.end()) {
                meshId = (int)fs->meshes.size();
                meshIds[mesh] = meshId;
                pb2_mesh m;
                std::memset(&m, 0, sizeof(m));
                m.first_tri = (int32_t)(fs->triIndex.size() / 3);
                m.n_tris = mesh->nTriangles;
                m.first_vertex = (int32_t)(fs->P.size() / 3);
                m.n_vertices = mesh->nVertices;
                m.has_n = !mesh->n.empty();
                m.has_uv = !mesh->uv.empty();
                m.has_s = !mesh->s.empty();
                m.reverse_orientation = tri->reverseOrientation;
                m.transform_swaps_handedness = tri->transformSwapsHandedness;
                m.alpha_tex = textureId(mesh->alphaMask);
                m.shadow_alpha_tex = textureId(mesh->shadowAlphaMask);
                anyN |= m.has_n != 0; anyUV |= m.has_uv != 0; anyS |= m.has_s != 0;
                for (int v = 0; v < mesh->nVertices; ++v) {
                    fs->P.insert(fs->P.end(), {mesh->p[v].x, mesh->p[v].y, mesh->p[v].z});
                    if (m.has_n) fs->N.insert(fs->N.end(), {mesh->n[v].x, mesh->n[v].y, mesh->n[v].z});
                    else fs->N.insert(fs->N.end(), {0.f, 0.f, 0.f});
                    if (m.has_uv) fs->UV.insert(fs->UV.end(), {mesh->uv[v].x, mesh->uv[v].y});
                    else fs->UV.insert(fs->UV.end(), {0.f, 0.f});
                    if (m.has_s) fs->S.insert(fs->S.end(), {mesh->s[v].x, mesh->s[v].y, mesh->s[v].z});
                    else fs->S.insert(fs->S.end(), {0.f, 0.f, 0.f});
                }
                for (int t = 0; t < mesh->nTriangles; ++t) {
                    for (int k = 0; k < 3; ++k) fs->triIndex.push_back(m.first_vertex + mesh->vertexIndices[3 * t + k]);
                    fs->triMesh.push_back(meshId);
                }
                fs->meshes.push_back(m);
            } else
                meshId = it->second;
            if (fs->meshes[meshId].reverse_orientation != (int)tri->reverseOrientation) {
                Error("Triangles of one mesh disagree on orientation");
                return nullptr;
            }
            fs->primType[i] = PB2_PRIM_TRIANGLE;
            fs->primIndex[i] = fs->meshes[meshId].first_tri + tri->triNumber;
        } else if (const Sphere *sp = dynamic_cast<const Sphere *>(gp->shape.get())) {
            pb2_sphere s;
            std::memset(&s, 0, sizeof(s));
            copyMatrix(sp->ObjectToWorld->GetMatrix(), s.object_to_world);
            copyMatrix(sp->WorldToObject->GetMatrix(), s.world_to_object);
            s.radius = sp->radius; s.z_min = sp->zMin; s.z_max = sp->zMax;
            s.theta_min = sp->thetaMin; s.theta_max = sp->thetaMax; s.phi_max = sp->phiMax;
            s.reverse_orientation = sp->reverseOrientation;
            s.transform_swaps_handedness = sp->transformSwapsHandedness;
            fs->primType[i] = PB2_PRIM_SPHERE;
            fs->primIndex[i] = (int32_t)fs->spheres.size();
            fs->spheres.push_back(s);
        } else {
            Error("Shape type outside the GPU path's scope (only triangles and spheres; SURVEY.md §2 row 27)");
            return nullptr;
        }
    }
    for (size_t i = 0; i < lights.size(); ++i)
        if (!lightSeen[i]) {
            Error("Area light %d is not attached to a primitive of the aggregate", (int)i);
            return nullptr;
        }

    pb2_scene_desc &d = fs->desc;
    std::memset(&d, 0, sizeof(d));
    d.n_vertices = (int64_t)(fs->P.size() / 3);
    d.P = fs->P.data();
    d.N = anyN ? fs->N.data() : nullptr;
    d.UV = anyUV ? fs->UV.data() : nullptr;
    d.S = anyS ? fs->S.data() : nullptr;
    d.n_tris = (int64_t)(fs->triIndex.size() / 3);
    d.tri_index = fs->triIndex.data();
    d.tri_mesh = fs->triMesh.data();
    d.n_meshes = (int32_t)fs->meshes.size();
    d.meshes = fs->meshes.data();
    d.n_spheres = (int32_t)fs->spheres.size();
    d.spheres = fs->spheres.data();
    d.n_prims = (int64_t)nPrims;
    d.prim_type = fs->primType.data();
    d.prim_index = fs->primIndex.data();
    d.prim_material = fs->primMaterial.data();
    d.prim_light = fs->primLight.data();
    d.n_nodes = (int64_t)bvh.nodes.size();
    d.nodes = bvh.nodes.data();
    d.bvh_prims = bvh.orderedPrimNumbers.data();
    if (!fs->instances.empty()) {
        // every BVH's nodes / ordered primitive numbers, scene BVH first; indices stay local to each BVH
        fs->nodes = bvh.nodes;
        fs->bvhPrims = bvh.orderedPrimNumbers;
        fs->bvhs.push_back(pb2_bvh{0, (int64_t)bvh.nodes.size(), 0, (int64_t)bvh.orderedPrimNumbers.size()});
        for (size_t k = 0; k < objectBvhs.size(); ++k) {
            const BVHAccel *ob = objectBvhs[k];
            fs->bvhs.push_back(pb2_bvh{(int64_t)fs->nodes.size(), (int64_t)ob->nodes.size(), (int64_t)fs->bvhPrims.size(),
                                       (int64_t)ob->orderedPrimNumbers.size()});
            fs->nodes.insert(fs->nodes.end(), ob->nodes.begin(), ob->nodes.end());
            for (int32_t n : ob->orderedPrimNumbers) fs->bvhPrims.push_back((int32_t)(objectBase[k] + n));
        }
        std::unordered_map<int32_t, int32_t> lonePosition;   // primitive number -> position in bvh_prims
        for (pb2_instance &inst : fs->instances) {
            if (inst.bvh >= 0) continue;
            auto it = lonePosition.find(inst.lone_prim);
            if (it == lonePosition.end()) {
                it = lonePosition.emplace(inst.lone_prim, (int32_t)fs->bvhPrims.size()).first;
                fs->bvhPrims.push_back(inst.lone_prim);
            }
            inst.lone_prim = it->second;
        }
        d.n_nodes = (int64_t)fs->nodes.size();
        d.nodes = fs->nodes.data();
        d.bvh_prims = fs->bvhPrims.data();
        d.n_bvh_prims = (int64_t)fs->bvhPrims.size();
        d.n_bvhs = (int32_t)fs->bvhs.size();
        d.bvhs = fs->bvhs.data();
        d.n_instances = (int32_t)fs->instances.size();
        d.instances = fs->instances.data();
    }
    d.n_materials = (int32_t)fs->materials.size();
    d.materials = fs->materials.data();
    for (const std::shared_ptr<ImageTexture> &t : fs->textureObjects) {
        pb2_texture pt;
        std::memset(&pt, 0, sizeof(pt));
        pt.channels = t->channels;
        pt.width = t->width;
        pt.height = t->height;
        pt.wrap = t->wrap;
        pt.do_trilinear = t->trilinear ? 1 : 0;
        pt.max_anisotropy = t->maxAniso;
        pt.su = t->su; pt.sv = t->sv; pt.du = t->du; pt.dv = t->dv;
        pt.texels = t->kind == PB2_TEXKIND_IMAGE ? t->texels.data() : nullptr;
        pt.kind = t->kind;
        for (int c = 0; c < 3; ++c) {
            pt.child[c] = t->child[c] ? textureIds[t->child[c].get()] : 0;
            pt.value[c] = t->value[c];
        }
        fs->textures.push_back(pt);
    }
    d.n_textures = (int32_t)fs->textures.size();
    d.textures = fs->textures.empty() ? nullptr : fs->textures.data();
    d.n_lights = (int32_t)fs->lights.size();
    d.lights = fs->lights.data();
    d.delta_lights = fs->deltaLights.empty() ? nullptr : fs->deltaLights.data();
    // lightdistrib.cpp:48-66: a single light always gets the uniform distribution
    if (lightStrategy == "uniform" || lights.size() == 1) d.light_strategy = PB2_LIGHTDIST_UNIFORM;
    else if (lightStrategy == "power") d.light_strategy = PB2_LIGHTDIST_POWER;
    else {
        if (lightStrategy != "spatial")
            Error("Light sample distribution type \"%s\" unknown. Using \"spatial\".", lightStrategy.c_str());
        d.light_strategy = PB2_LIGHTDIST_SPATIAL;
    }
    d.spatial_max_voxels = 64;
    return fs;
}

}  // namespace pbrt
