// Four-child node records for the trace kernel: groundwork for the next kernel generation, NOT used by any
// render path yet.  The record builder and the visit function below are exercised on the host by
// tests/test_host.py::test_wide4_records_keep_the_reference_order through pb2_debug_wide4_sequences.
//
// One 128-B record per interior node of every second level of the reference's binary tree (src/accelerators/bvh.cpp:
// 95-104, 640-658): the boxes of the node's grandchildren in the canonical slot order [LL, LR, RL, RR]; a child that is a
// leaf keeps its own box in the first slot of its pair and leaves the second empty.
//   q[0..5]  24 floats: slot s has min xyz at f[6 s .. 6 s + 2], max xyz at f[6 s + 3 .. 6 s + 5]
//   q[6]     the four child references (int bits): index of the child's record, a WIDE_LEAF reference as in the two-child
//            records (pb2_scene.cuh), or WIDE4_EMPTY
//   q[7].x   meta (int bits): split axis of the node (bits 0-1), of its first child (2-3), of its second child (4-5)
// Visiting order = the reference's (bvh.cpp:682-690): at each of the two collapsed levels the child on the ray's near side
// of the split axis first.  A child box is tested with the tMax of the visit; the caller re-checks a deferred child's
// entry distance against the tMax of the moment it is taken up.  Since a child's slab interval lies inside its parent's
// and tMax only shrinks, every leaf is reached iff the binary traversal reaches it, in the same order.
#ifndef PB2_WIDE4_CUH
#define PB2_WIDE4_CUH

#include <cstring>
#include <vector>

#include "pb2_scene.cuh"

namespace pb2 {

enum : uint32_t { WIDE4_EMPTY = 0x7fffffffu };

// Host: collapse nodes[0 .. nNodes) into four-child records.  Record 0 belongs to the root; a tree that is a single leaf
// gets one record whose first slot is that leaf.
inline std::vector<float4> buildWide4Records(const pb2_bvh_node *nodes, int64_t nNodes) {
    std::vector<float4> out;
    if (nNodes <= 0) return out;
    std::vector<int32_t> recordOf((size_t)nNodes, -1);   // binary interior node -> its record
    std::vector<int32_t> order;                           // binary nodes that own a record, in record order
    auto recordFor = [&](int32_t node) {
        if (recordOf[node] < 0) {
            recordOf[node] = (int32_t)order.size();
            order.push_back(node);
        }
        return recordOf[node];
    };
    auto leafRef = [&](const pb2_bvh_node &n) { return WIDE_LEAF | ((uint32_t)(n.n_prims - 1) << WIDE_LEAF_COUNT_SHIFT) | (uint32_t)n.offset; };
    recordFor(0);
    for (size_t k = 0; k < order.size(); ++k) {   // order grows while it is walked
        const int32_t i = order[k];
        float f[32];
        uint32_t refs[4] = {WIDE4_EMPTY, WIDE4_EMPTY, WIDE4_EMPTY, WIDE4_EMPTY};
        for (int s = 0; s < 4; ++s)
            for (int c = 0; c < 6; ++c) f[6 * s + c] = c < 3 ? PB2_INFINITY : -PB2_INFINITY;
        uint32_t meta = 0;
        auto put = [&](int slot, int32_t node) {
            const pb2_bvh_node &n = nodes[node];
            for (int c = 0; c < 3; ++c) {
                f[6 * slot + c] = n.bmin[c];
                f[6 * slot + 3 + c] = n.bmax[c];
            }
            refs[slot] = n.n_prims > 0 ? leafRef(n) : (uint32_t)recordFor(node);
        };
        if (nodes[i].n_prims > 0)
            put(0, i);   // the whole tree is one leaf
        else {
            meta = nodes[i].axis;
            const int32_t child[2] = {i + 1, nodes[i].offset};
            for (int g = 0; g < 2; ++g) {
                const pb2_bvh_node &c = nodes[child[g]];
                if (c.n_prims > 0)
                    put(2 * g, child[g]);
                else {
                    meta |= (uint32_t)c.axis << (2 + 2 * g);
                    put(2 * g, child[g] + 1);
                    put(2 * g + 1, c.offset);
                }
            }
        }
        const size_t base = out.size();
        out.resize(base + 8);
        std::memcpy(&out[base], f, 24 * sizeof(float));
        std::memcpy(&out[base + 6], refs, sizeof(refs));
        std::memcpy(&out[base + 7].x, &meta, sizeof(meta));
    }
    return out;
}

// One visit: tests the record's children against the ray with the current tMax and returns those that pass, nearest in
// the reference's visiting order first, with their entry distances.  The first is the one to continue with; the others are
// deferred in REVERSE order (last pushed = next visited) and re-checked (tMin < tMax) when popped.
PB2_HD int wide4Visit(const float4 *rec, const DRaySetup &r, float tMax, uint32_t refs[4], float tMins[4]) {
    const float *f = reinterpret_cast<const float *>(rec);
    const uint32_t *ref = reinterpret_cast<const uint32_t *>(rec + 6);
    const uint32_t meta = reinterpret_cast<const uint32_t *>(rec + 7)[0];
    const int neg[3] = {r.neg0, r.neg1, r.neg2};
    const int negTop = neg[meta & 3], negPair[2] = {neg[(meta >> 2) & 3], neg[(meta >> 4) & 3]};
    int n = 0;
    for (int p = 0; p < 4; ++p) {
        const int g = (p >> 1) ^ negTop;          // which child of the node comes first
        const int s = 2 * g + ((p & 1) ^ negPair[g]);   // which of its children
        if (ref[s] == WIDE4_EMPTY) continue;
        float tMin;
        if (slabTestT(f[6 * s], f[6 * s + 1], f[6 * s + 2], f[6 * s + 3], f[6 * s + 4], f[6 * s + 5], r, tMax, &tMin)) {
            refs[n] = ref[s];
            tMins[n] = tMin;
            ++n;
        }
    }
    return n;
}

}  // namespace pb2
#endif
