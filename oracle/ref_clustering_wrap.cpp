// ref_clustering_wrap.cpp -- C wrapper around the REFERENCE's own graph clustering, compiled from
// the reference sources where they lie (/root/reference/src/clustering.cc, include/clustering.h,
// include/universe.h: STL only) into oracle/_ref/libref_clustering.so by `make -C oracle ref`.
// Test infrastructure: it pins the oracle's restatement (orc_kat_cluster) and the product's host
// clustering (l3d_cluster_edges) against the reference itself.  Nothing of the reference is copied:
// this file only calls L3DPP::performClustering (src/clustering.cc:7-48) and reads the result the way
// Line3D::clusterSegments does (CLUniverse::find per local id, src/line3D.cc:2522-2535).
#include <list>

#include "clustering.h"  // -I/root/reference/include

extern "C" int ref_cluster(const int* ij, const float* w, int ne, int n, float c, int* out)
{
    std::list<L3DPP::CLEdge> edges;
    for (int e = 0; e < ne; ++e) {
        L3DPP::CLEdge ed;
        ed.i_ = ij[2 * e];
        ed.j_ = ij[2 * e + 1];
        ed.w_ = w[e];
        edges.push_back(ed);
    }
    L3DPP::CLUniverse* u = L3DPP::performClustering(edges, n, c);
    if (!u) return 0;
    for (int i = 0; i < n; ++i) out[i] = u->find(i);
    delete u;
    return n;
}
