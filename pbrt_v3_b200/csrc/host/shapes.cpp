// Host-side shape producers: everything that ends in a world-space TriangleMesh (trianglemesh,
// loopsubdiv, plymesh) plus the Sphere record.  Scene preparation only — ray intersection for these
// shapes happens on the GPU (../device/intersect.cuh).
//
//   TriangleMesh ctor / CreateTriangleMesh     src/shapes/triangle.cpp:54-110
//   CreateTriangleMeshShape                    src/shapes/triangle.cpp:647-743
//   Loop subdivision                           src/shapes/loopsubdiv.cpp:155-470
//   PLY meshes                                 src/shapes/plymesh.cpp
//   Sphere ctor / bounds / factory             src/shapes/sphere.h:50-59, sphere.cpp:44-47,329-340
#include <map>
#include <set>

#include "scene.h"

namespace pbrt {

// ---------------------------------------------------------------- triangle meshes
TriangleMesh::TriangleMesh(const Transform &ObjectToWorld, int nTriangles, const int *vertexIndices, int nVertices,
                           const Point3f *P, const Vector3f *S, const Normal3f *N, const Point2f *UV)
    : nTriangles(nTriangles), nVertices(nVertices), vertexIndices(vertexIndices, vertexIndices + 3 * nTriangles) {
    p.resize(nVertices);
    for (int i = 0; i < nVertices; ++i) p[i] = ObjectToWorld(P[i]);
    if (UV) uv.assign(UV, UV + nVertices);
    if (N) {
        n.resize(nVertices);
        for (int i = 0; i < nVertices; ++i) n[i] = ObjectToWorld.ApplyNormal(N[i]);
    }
    if (S) {
        s.resize(nVertices);
        for (int i = 0; i < nVertices; ++i) s[i] = ObjectToWorld.ApplyVector(S[i]);
    }
}

Bounds3f Triangle::ObjectBound() const {
    const Point3f &p0 = mesh->p[v[0]], &p1 = mesh->p[v[1]], &p2 = mesh->p[v[2]];
    return Union(Bounds3f((*WorldToObject)(p0), (*WorldToObject)(p1)), (*WorldToObject)(p2));
}
Bounds3f Triangle::WorldBound() const {
    const Point3f &p0 = mesh->p[v[0]], &p1 = mesh->p[v[1]], &p2 = mesh->p[v[2]];
    return Union(Bounds3f(p0, p1), p2);
}
Float Triangle::Area() const {
    const Point3f &p0 = mesh->p[v[0]], &p1 = mesh->p[v[1]], &p2 = mesh->p[v[2]];
    return 0.5 * Cross(p1 - p0, p2 - p0).Length();
}

std::vector<std::shared_ptr<Shape>> CreateTriangleMesh(const Transform *o2w, const Transform *w2o, bool reverseOrientation,
                                                       int nTriangles, const int *vertexIndices, int nVertices,
                                                       const Point3f *p, const Vector3f *s, const Normal3f *n,
                                                       const Point2f *uv) {
    auto mesh = std::make_shared<TriangleMesh>(*o2w, nTriangles, vertexIndices, nVertices, p, s, n, uv);
    std::vector<std::shared_ptr<Shape>> tris;
    tris.reserve(nTriangles);
    for (int i = 0; i < nTriangles; ++i)
        tris.push_back(std::make_shared<Triangle>(o2w, w2o, reverseOrientation, mesh, i));
    return tris;
}

static std::vector<Point3f> toPoints(const std::vector<Float> &f) {
    std::vector<Point3f> r(f.size() / 3);
    for (size_t i = 0; i < r.size(); ++i) r[i] = Point3f(f[3 * i], f[3 * i + 1], f[3 * i + 2]);
    return r;
}

std::vector<std::shared_ptr<Shape>> CreateTriangleMeshShape(const Transform *o2w, const Transform *w2o,
                                                            bool reverseOrientation, const ParamSet &params) {
    typedef ParamSet::Type T;
    bool haveVi = false, haveP = false;
    std::vector<int> vi = params.FindInts("indices", &haveVi);
    std::vector<Point3f> P = toPoints(params.FindFloats(T::Point3, "P", &haveP));
    int npi = (int)P.size();
    // "uv"/"st" may be given as point2 or as plain floats (triangle.cpp:654-667)
    bool haveUV = false;
    std::vector<Float> fuv = params.FindFloats(T::Point2, "uv", &haveUV);
    if (!haveUV) fuv = params.FindFloats(T::Point2, "st", &haveUV);
    if (!haveUV) fuv = params.FindFloats(T::Float, "uv", &haveUV);
    if (!haveUV) fuv = params.FindFloats(T::Float, "st", &haveUV);
    std::vector<Point2f> uvs;
    if (haveUV) {
        int nuvi = (int)fuv.size() / 2;
        if (nuvi < npi) {
            Error("Not enough of \"uv\"s for triangle mesh.  Expected %d, found %d.  Discarding.", npi, nuvi);
        } else {
            if (nuvi > npi)
                Warning("More \"uv\"s provided than will be used for triangle mesh.  (%d expcted, %d found)", npi, nuvi);
            uvs.resize(nuvi);
            for (int i = 0; i < nuvi; ++i) uvs[i] = Point2f(fuv[2 * i], fuv[2 * i + 1]);
        }
    }
    if (!haveVi) {
        Error("Vertex indices \"indices\" not provided with triangle mesh shape");
        return {};
    }
    if (!haveP) {
        Error("Vertex positions \"P\" not provided with triangle mesh shape");
        return {};
    }
    bool haveS = false, haveN = false;
    std::vector<Vector3f> S = toPoints(params.FindFloats(T::Vector3, "S", &haveS));
    if (haveS && (int)S.size() != npi) {
        Error("Number of \"S\"s for triangle mesh must match \"P\"s");
        haveS = false;
    }
    std::vector<Normal3f> N = toPoints(params.FindFloats(T::Normal, "N", &haveN));
    if (haveN && (int)N.size() != npi) {
        Error("Number of \"N\"s for triangle mesh must match \"P\"s");
        haveN = false;
    }
    for (size_t i = 0; i < vi.size(); ++i)
        if (vi[i] >= npi) {
            Error("trianglemesh has out of-bounds vertex index %d (%d \"P\" values were given", vi[i], npi);
            return {};
        }
    params.FindInts("faceIndices");
    return CreateTriangleMesh(o2w, w2o, reverseOrientation, (int)vi.size() / 3, vi.data(), npi, P.data(),
                              haveS ? S.data() : nullptr, haveN ? N.data() : nullptr,
                              uvs.empty() ? nullptr : uvs.data());
}

// ---------------------------------------------------------------- Loop subdivision
// Index-based half-edge-free formulation of the reference's pointer mesh: faces hold three vertex
// ids and three neighbour-face ids (neighbour k shares edge v[k]→v[k+1]).  Iteration orders (faces
// in order, edges k=0..2, new vertices appended as first encountered) are the reference's, which
// fixes both the output vertex numbering and the float summation order.
namespace {
struct SVert {
    Point3f p;
    int startFace = -1;
    int child = -1;
    bool regular = false, boundary = false;
};
struct SFace {
    int v[3] = {-1, -1, -1};
    int f[3] = {-1, -1, -1};
    int children[4] = {-1, -1, -1, -1};
};
inline int nxt(int i) { return (i + 1) % 3; }
inline int prv(int i) { return (i + 2) % 3; }

struct SubdivMesh {
    std::vector<SVert> V;
    std::vector<SFace> F;
    int vnum(int f, int vert) const {
        for (int i = 0; i < 3; ++i)
            if (F[f].v[i] == vert) return i;
        Error("Basic logic error in loop subdivision vnum()");
        return 0;
    }
    int nextFace(int f, int vert) const { return F[f].f[vnum(f, vert)]; }
    int prevFace(int f, int vert) const { return F[f].f[prv(vnum(f, vert))]; }
    int nextVert(int f, int vert) const { return F[f].v[nxt(vnum(f, vert))]; }
    int prevVert(int f, int vert) const { return F[f].v[prv(vnum(f, vert))]; }
    int otherVert(int f, int v0, int v1) const {
        for (int i = 0; i < 3; ++i)
            if (F[f].v[i] != v0 && F[f].v[i] != v1) return F[f].v[i];
        Error("Basic logic error in loop subdivision otherVert()");
        return 0;
    }
    int valence(int vert) const {
        int f = V[vert].startFace;
        if (!V[vert].boundary) {
            int nf = 1;
            while ((f = nextFace(f, vert)) != V[vert].startFace) ++nf;
            return nf;
        }
        int nf = 1;
        while ((f = nextFace(f, vert)) != -1) ++nf;
        f = V[vert].startFace;
        while ((f = prevFace(f, vert)) != -1) ++nf;
        return nf + 1;
    }
    void oneRing(int vert, Point3f *out) const {
        if (!V[vert].boundary) {
            int face = V[vert].startFace;
            do {
                *out++ = V[nextVert(face, vert)].p;
                face = nextFace(face, vert);
            } while (face != V[vert].startFace);
        } else {
            int face = V[vert].startFace, f2;
            while ((f2 = nextFace(face, vert)) != -1) face = f2;
            *out++ = V[nextVert(face, vert)].p;
            do {
                *out++ = V[prevVert(face, vert)].p;
                face = prevFace(face, vert);
            } while (face != -1);
        }
    }
    Point3f weightOneRing(int vert, Float beta) const {
        int val = valence(vert);
        std::vector<Point3f> ring(val);
        oneRing(vert, ring.data());
        Point3f p = (1 - val * beta) * V[vert].p;
        for (int i = 0; i < val; ++i) p = p + beta * ring[i];
        return p;
    }
    Point3f weightBoundary(int vert, Float beta) const {
        int val = valence(vert);
        std::vector<Point3f> ring(val);
        oneRing(vert, ring.data());
        Point3f p = (1 - 2 * beta) * V[vert].p;
        p = p + beta * ring[0];
        p = p + beta * ring[val - 1];
        return p;
    }
};
inline Float loopBeta(int valence) { return valence == 3 ? 3.f / 16.f : 3.f / (8.f * valence); }
inline Float loopGamma(int valence) { return 1.f / (valence + 3.f / (8.f * loopBeta(valence))); }
typedef std::pair<int, int> EdgeKey;
inline EdgeKey edgeKey(int a, int b) { return EdgeKey(std::min(a, b), std::max(a, b)); }
}  // namespace

void LoopSubdivide(int nLevels, int nIndices, const int *vertexIndices, int nVertices, const Point3f *p,
                   std::vector<Point3f> *pLimitOut, std::vector<Normal3f> *NsOut, std::vector<int> *indicesOut) {
    SubdivMesh M;
    M.V.resize(nVertices);
    for (int i = 0; i < nVertices; ++i) M.V[i].p = p[i];
    int nFaces = nIndices / 3;
    M.F.resize(nFaces);
    for (int i = 0; i < nFaces; ++i)
        for (int j = 0; j < 3; ++j) {
            int v = vertexIndices[3 * i + j];
            M.F[i].v[j] = v;
            M.V[v].startFace = i;
        }
    // neighbour pointers (loopsubdiv.cpp:177-197): first face to see an edge waits for the second
    {
        std::map<EdgeKey, std::pair<int, int>> open;  // edge -> (face, edgeNum)
        for (int i = 0; i < nFaces; ++i)
            for (int e = 0; e < 3; ++e) {
                EdgeKey k = edgeKey(M.F[i].v[e], M.F[i].v[nxt(e)]);
                auto it = open.find(k);
                if (it == open.end())
                    open[k] = std::make_pair(i, e);
                else {
                    M.F[it->second.first].f[it->second.second] = i;
                    M.F[i].f[e] = it->second.first;
                    open.erase(it);
                }
            }
    }
    for (int i = 0; i < nVertices; ++i) {
        int f = M.V[i].startFace;
        do {
            f = M.nextFace(f, i);
        } while (f != -1 && f != M.V[i].startFace);
        M.V[i].boundary = (f == -1);
        int val = M.valence(i);
        M.V[i].regular = (!M.V[i].boundary && val == 6) || (M.V[i].boundary && val == 4);
    }

    // current level = index lists into the growing V/F pools
    std::vector<int> fcur(nFaces), vcur(nVertices);
    for (int i = 0; i < nFaces; ++i) fcur[i] = i;
    for (int i = 0; i < nVertices; ++i) vcur[i] = i;
    for (int level = 0; level < nLevels; ++level) {
        std::vector<int> newF, newV;
        for (int v : vcur) {
            SVert c;
            c.regular = M.V[v].regular;
            c.boundary = M.V[v].boundary;
            M.V.push_back(c);
            M.V[v].child = (int)M.V.size() - 1;
            newV.push_back(M.V[v].child);
        }
        for (int f : fcur)
            for (int k = 0; k < 4; ++k) {
                M.F.push_back(SFace());
                M.F[f].children[k] = (int)M.F.size() - 1;
                newF.push_back(M.F[f].children[k]);
            }
        // even vertices
        for (int v : vcur) {
            Point3f np;
            if (!M.V[v].boundary)
                np = M.V[v].regular ? M.weightOneRing(v, 1.f / 16.f) : M.weightOneRing(v, loopBeta(M.valence(v)));
            else
                np = M.weightBoundary(v, 1.f / 8.f);
            M.V[M.V[v].child].p = np;
        }
        // odd (edge) vertices
        std::map<EdgeKey, int> edgeVerts;
        for (int f : fcur)
            for (int k = 0; k < 3; ++k) {
                int a = M.F[f].v[k], b = M.F[f].v[nxt(k)];
                EdgeKey key = edgeKey(a, b);
                if (edgeVerts.count(key)) continue;
                SVert nv;
                nv.regular = true;
                nv.boundary = (M.F[f].f[k] == -1);
                nv.startFace = M.F[f].children[3];
                // key.first/second play the role of SDEdge::v[0]/v[1]; the two leading terms commute
                if (nv.boundary) {
                    nv.p = 0.5f * M.V[key.first].p;
                    nv.p = nv.p + 0.5f * M.V[key.second].p;
                } else {
                    nv.p = 3.f / 8.f * M.V[key.first].p;
                    nv.p = nv.p + 3.f / 8.f * M.V[key.second].p;
                    nv.p = nv.p + 1.f / 8.f * M.V[M.otherVert(f, a, b)].p;
                    nv.p = nv.p + 1.f / 8.f * M.V[M.otherVert(M.F[f].f[k], a, b)].p;
                }
                M.V.push_back(nv);
                edgeVerts[key] = (int)M.V.size() - 1;
                newV.push_back((int)M.V.size() - 1);
            }
        // topology of the next level
        for (int v : vcur) {
            int vertNum = M.vnum(M.V[v].startFace, v);
            M.V[M.V[v].child].startFace = M.F[M.V[v].startFace].children[vertNum];
        }
        for (int f : fcur)
            for (int j = 0; j < 3; ++j) {
                const int *ch = M.F[f].children;
                M.F[ch[3]].f[j] = ch[nxt(j)];
                M.F[ch[j]].f[nxt(j)] = ch[3];
                int f2 = M.F[f].f[j];
                M.F[ch[j]].f[j] = (f2 != -1) ? M.F[f2].children[M.vnum(f2, M.F[f].v[j])] : -1;
                f2 = M.F[f].f[prv(j)];
                M.F[ch[j]].f[prv(j)] = (f2 != -1) ? M.F[f2].children[M.vnum(f2, M.F[f].v[j])] : -1;
            }
        for (int f : fcur)
            for (int j = 0; j < 3; ++j) {
                const int *ch = M.F[f].children;
                M.F[ch[j]].v[j] = M.V[M.F[f].v[j]].child;
                int vert = edgeVerts[edgeKey(M.F[f].v[j], M.F[f].v[nxt(j)])];
                M.F[ch[j]].v[nxt(j)] = vert;
                M.F[ch[nxt(j)]].v[j] = vert;
                M.F[ch[3]].v[j] = vert;
            }
        fcur.swap(newF);
        vcur.swap(newV);
    }

    // limit surface positions
    std::vector<Point3f> pLimit(vcur.size());
    for (size_t i = 0; i < vcur.size(); ++i) {
        int v = vcur[i];
        pLimit[i] = M.V[v].boundary ? M.weightBoundary(v, 1.f / 5.f) : M.weightOneRing(v, loopGamma(M.valence(v)));
    }
    for (size_t i = 0; i < vcur.size(); ++i) M.V[vcur[i]].p = pLimit[i];

    // limit surface normals from the tangent masks
    std::vector<Normal3f> Ns;
    Ns.reserve(vcur.size());
    std::vector<Point3f> ring(16);
    for (int v : vcur) {
        Vector3f S(0, 0, 0), T(0, 0, 0);
        int val = M.valence(v);
        if (val > (int)ring.size()) ring.resize(val);
        M.oneRing(v, ring.data());
        const Point3f &vp = M.V[v].p;
        if (!M.V[v].boundary) {
            for (int j = 0; j < val; ++j) {
                S = S + std::cos(2 * Pi * j / val) * ring[j];
                T = T + std::sin(2 * Pi * j / val) * ring[j];
            }
        } else {
            S = ring[val - 1] - ring[0];
            if (val == 2)
                T = ring[0] + ring[1] - 2 * vp;
            else if (val == 3)
                T = ring[1] - vp;
            else if (val == 4)
                T = -1 * ring[0] + 2 * ring[1] + 2 * ring[2] + -1 * ring[3] + -2 * vp;
            else {
                Float theta = Pi / float(val - 1);
                T = std::sin(theta) * (ring[0] + ring[val - 1]);
                for (int k = 1; k < val - 1; ++k) {
                    Float wt = (2 * std::cos(theta) - 2) * std::sin((k)*theta);
                    T = T + wt * ring[k];
                }
                T = -T;
            }
        }
        Ns.push_back(Cross(S, T));
    }

    std::map<int, int> used;
    for (size_t i = 0; i < vcur.size(); ++i) used[vcur[i]] = (int)i;
    indicesOut->clear();
    indicesOut->reserve(3 * fcur.size());
    for (int f : fcur)
        for (int j = 0; j < 3; ++j) indicesOut->push_back(used[M.F[f].v[j]]);
    *pLimitOut = std::move(pLimit);
    *NsOut = std::move(Ns);
}

std::vector<std::shared_ptr<Shape>> CreateLoopSubdiv(const Transform *o2w, const Transform *w2o,
                                                     bool reverseOrientation, const ParamSet &params) {
    int nLevels = params.FindOneInt("levels", params.FindOneInt("nlevels", 3));
    bool haveVi = false, haveP = false;
    std::vector<int> vi = params.FindInts("indices", &haveVi);
    std::vector<Point3f> P = toPoints(params.FindFloats(ParamSet::Type::Point3, "P", &haveP));
    if (!haveVi) {
        Error("Vertex indices \"indices\" not provided for LoopSubdiv shape.");
        return {};
    }
    if (!haveP) {
        Error("Vertex positions \"P\" not provided for LoopSubdiv shape.");
        return {};
    }
    params.FindOneString("scheme", "loop");
    std::vector<Point3f> pLimit;
    std::vector<Normal3f> Ns;
    std::vector<int> idx;
    LoopSubdivide(nLevels, (int)vi.size(), vi.data(), (int)P.size(), P.data(), &pLimit, &Ns, &idx);
    return CreateTriangleMesh(o2w, w2o, reverseOrientation, (int)idx.size() / 3, idx.data(), (int)pLimit.size(),
                              pLimit.data(), nullptr, Ns.data(), nullptr);
}

// ---------------------------------------------------------------- PLY
namespace {
struct PlyProp {
    std::string name;
    std::string type, countType;  // countType non-empty for list properties
};
int plyTypeSize(const std::string &t) {
    if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
    if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
    if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
    if (t == "double" || t == "float64") return 8;
    return 0;
}
double plyReadBinary(const unsigned char *&src, const std::string &t, bool bigEndian) {
    double v = 0;
    unsigned char swapped[8];
    const unsigned char *ptr = src;
    const int size = plyTypeSize(t);
    if (bigEndian && size > 1) {
        for (int i = 0; i < size; ++i) swapped[i] = src[size - 1 - i];
        ptr = swapped;
    }
    if (t == "char" || t == "int8") v = *(const int8_t *)ptr;
    else if (t == "uchar" || t == "uint8") v = *(const uint8_t *)ptr;
    else if (t == "short" || t == "int16") { int16_t x; std::memcpy(&x, ptr, 2); v = x; }
    else if (t == "ushort" || t == "uint16") { uint16_t x; std::memcpy(&x, ptr, 2); v = x; }
    else if (t == "int" || t == "int32") { int32_t x; std::memcpy(&x, ptr, 4); v = x; }
    else if (t == "uint" || t == "uint32") { uint32_t x; std::memcpy(&x, ptr, 4); v = x; }
    else if (t == "float" || t == "float32") { float x; std::memcpy(&x, ptr, 4); v = x; }
    else if (t == "double" || t == "float64") { double x; std::memcpy(&x, ptr, 8); v = x; }
    src += size;
    return v;
}
}  // namespace

bool ReadPLY(const std::string &filename, std::vector<Point3f> *P, std::vector<Normal3f> *N,
             std::vector<Point2f> *UV, std::vector<int> *indices) {
    FILE *fp = std::fopen(filename.c_str(), "rb");
    if (!fp) {
        Error("Couldn't open PLY file \"%s\"", filename.c_str());
        return false;
    }
    std::vector<unsigned char> data;
    {
        std::fseek(fp, 0, SEEK_END);
        long sz = std::ftell(fp);
        std::fseek(fp, 0, SEEK_SET);
        data.resize(sz);
        if (sz && std::fread(data.data(), 1, sz, fp) != (size_t)sz) {
            std::fclose(fp);
            Error("Short read on PLY file \"%s\"", filename.c_str());
            return false;
        }
        std::fclose(fp);
    }
    // header
    size_t pos = 0;
    auto readLine = [&]() {
        std::string line;
        while (pos < data.size() && data[pos] != '\n') line.push_back((char)data[pos++]);
        if (pos < data.size()) ++pos;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        return line;
    };
    if (readLine() != "ply") {
        Error("\"%s\" is not a PLY file", filename.c_str());
        return false;
    }
    bool ascii = false, bigEndian = false;
    struct Element { std::string name; long count; std::vector<PlyProp> props; };
    std::vector<Element> elems;
    while (pos < data.size()) {
        std::string line = readLine();
        char a[64], b[64], c[64], d[64];
        if (line == "end_header") break;
        if (std::sscanf(line.c_str(), "format %63s", a) == 1) {
            std::string fmt(a);
            if (fmt == "ascii") ascii = true;
            else if (fmt == "binary_big_endian") bigEndian = true;
            else if (fmt != "binary_little_endian") {
                Error("PLY format \"%s\" not supported", a);
                return false;
            }
        } else if (std::sscanf(line.c_str(), "element %63s %63s", a, b) == 2) {
            elems.push_back(Element{a, std::atol(b), {}});
        } else if (std::sscanf(line.c_str(), "property list %63s %63s %63s", a, b, c) == 3) {
            if (!elems.empty()) elems.back().props.push_back(PlyProp{c, b, a});
        } else if (std::sscanf(line.c_str(), "property %63s %63s", a, d) == 2) {
            if (!elems.empty()) elems.back().props.push_back(PlyProp{d, a, ""});
        }
    }
    const unsigned char *ptr = data.data() + pos;
    const char *txt = (const char *)ptr;
    auto nextAscii = [&]() {
        char *end = nullptr;
        double v = std::strtod(txt, &end);
        txt = end;
        return v;
    };
    for (const Element &e : elems) {
        bool isVertex = e.name == "vertex", isFace = e.name == "face";
        int ix = -1, iy = -1, iz = -1, inx = -1, iny = -1, inz = -1, iu = -1, iv = -1;
        for (size_t k = 0; k < e.props.size(); ++k) {
            const std::string &n = e.props[k].name;
            if (n == "x") ix = (int)k; else if (n == "y") iy = (int)k; else if (n == "z") iz = (int)k;
            else if (n == "nx") inx = (int)k; else if (n == "ny") iny = (int)k; else if (n == "nz") inz = (int)k;
            else if (n == "u" || n == "s" || n == "texture_u" || n == "texture_s") iu = (int)k;
            else if (n == "v" || n == "t" || n == "texture_v" || n == "texture_t") iv = (int)k;
        }
        bool hasN = inx >= 0 && iny >= 0 && inz >= 0, hasUV = iu >= 0 && iv >= 0;
        if (isVertex) {
            P->resize(e.count);
            if (hasN) N->resize(e.count);
            if (hasUV) UV->resize(e.count);
        }
        std::vector<double> vals(e.props.size());
        for (long i = 0; i < e.count; ++i) {
            for (size_t k = 0; k < e.props.size(); ++k) {
                const PlyProp &pr = e.props[k];
                if (pr.countType.empty()) {
                    vals[k] = ascii ? nextAscii() : plyReadBinary(ptr, pr.type, bigEndian);
                } else {
                    int cnt = (int)(ascii ? nextAscii() : plyReadBinary(ptr, pr.countType, bigEndian));
                    std::vector<int> lst(cnt);
                    for (int j = 0; j < cnt; ++j) lst[j] = (int)(ascii ? nextAscii() : plyReadBinary(ptr, pr.type, bigEndian));
                    if (isFace && (pr.name == "vertex_indices" || pr.name == "vertex_index")) {
                        // plymesh.cpp:123-156: triangles as-is, quads split (0,1,2) (2,3,0)... via a fan
                        if (cnt == 3) {
                            indices->insert(indices->end(), {lst[0], lst[1], lst[2]});
                        } else if (cnt == 4) {
                            indices->insert(indices->end(), {lst[0], lst[1], lst[2], lst[3], lst[0], lst[2]});
                        } else {
                            // plymesh.cpp:112-116
                            Warning("plymesh: Ignoring face with %i vertices (only triangles and quads are supported!)", cnt);
                        }
                    }
                }
            }
            if (isVertex) {
                (*P)[i] = Point3f((Float)vals[ix], (Float)vals[iy], (Float)vals[iz]);
                if (hasN) (*N)[i] = Normal3f((Float)vals[inx], (Float)vals[iny], (Float)vals[inz]);
                if (hasUV) (*UV)[i] = Point2f((Float)vals[iu], (Float)vals[iv]);
            }
        }
    }
    return true;
}

bool WritePLY(const std::string &filename, const std::vector<Point3f> &P, const std::vector<int> &indices) {
    FILE *fp = std::fopen(filename.c_str(), "wb");
    if (!fp) return false;
    std::fprintf(fp, "ply\nformat binary_little_endian 1.0\nelement vertex %zu\nproperty float x\nproperty float y\n"
                     "property float z\nelement face %zu\nproperty list uchar int vertex_indices\nend_header\n",
                 P.size(), indices.size() / 3);
    for (const Point3f &p : P) {
        float v[3] = {p.x, p.y, p.z};
        std::fwrite(v, 4, 3, fp);
    }
    for (size_t i = 0; i + 2 < indices.size(); i += 3) {
        unsigned char c = 3;
        std::fwrite(&c, 1, 1, fp);
        int32_t v[3] = {indices[i], indices[i + 1], indices[i + 2]};
        std::fwrite(v, 4, 3, fp);
    }
    std::fclose(fp);
    return true;
}

extern std::string g_sceneDirectory;  // api.cpp: directory of the file being parsed
std::vector<std::shared_ptr<Shape>> CreatePLYMesh(const Transform *o2w, const Transform *w2o,
                                                  bool reverseOrientation, const ParamSet &params) {
    std::string filename = params.FindOneString("filename", "");
    if (!filename.empty() && filename[0] != '/' && !g_sceneDirectory.empty()) filename = g_sceneDirectory + "/" + filename;
    std::vector<Point3f> P;
    std::vector<Normal3f> N;
    std::vector<Point2f> UV;
    std::vector<int> idx;
    if (!ReadPLY(filename, &P, &N, &UV, &idx)) return {};
    if (P.empty() || idx.empty()) {
        Error("PLY file \"%s\" is invalid! No face/vertex elements found!", filename.c_str());
        return {};
    }
    return CreateTriangleMesh(o2w, w2o, reverseOrientation, (int)idx.size() / 3, idx.data(), (int)P.size(), P.data(),
                              nullptr, N.empty() ? nullptr : N.data(), UV.empty() ? nullptr : UV.data());
}

// ---------------------------------------------------------------- sphere
Sphere::Sphere(const Transform *o2w, const Transform *w2o, bool reverseOrientation, Float radius, Float zMin,
               Float zMax, Float phiMax)
    : Shape(o2w, w2o, reverseOrientation),
      radius(radius),
      zMin(Clamp(std::min(zMin, zMax), -radius, radius)),
      zMax(Clamp(std::max(zMin, zMax), -radius, radius)),
      thetaMin(std::acos(Clamp(std::min(zMin, zMax) / radius, -1, 1))),
      thetaMax(std::acos(Clamp(std::max(zMin, zMax) / radius, -1, 1))),
      phiMax(Radians(Clamp(phiMax, 0, 360))) {}

Bounds3f Sphere::ObjectBound() const { return Bounds3f(Point3f(-radius, -radius, zMin), Point3f(radius, radius, zMax)); }

std::shared_ptr<Shape> CreateSphereShape(const Transform *o2w, const Transform *w2o, bool reverseOrientation,
                                         const ParamSet &params) {
    Float radius = params.FindOneFloat("radius", 1.f);
    Float zmin = params.FindOneFloat("zmin", -radius);
    Float zmax = params.FindOneFloat("zmax", radius);
    Float phimax = params.FindOneFloat("phimax", 360.f);
    return std::make_shared<Sphere>(o2w, w2o, reverseOrientation, radius, zmin, zmax, phimax);
}

}  // namespace pbrt
