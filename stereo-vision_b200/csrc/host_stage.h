// Host middle stage of the pipeline: the parts of Elas::process that are inherently sequential
// and tiny (tens of kilobytes), executed on the CPU between the two GPU phases of a frame:
//
//   candidate lattice (from K2)  ->  in-place lattice filters      elas.cpp:174-279, :496-502
//                                ->  support point list            elas.cpp:505-523
//                                ->  Delaunay triangulation x2     elas.cpp:534-600 (Triangle "zQB")
//                                ->  disparity planes              elas.cpp:605-680 (Matrix::solve)
//                                ->  per-triangle raster records   elas.cpp:1006-1072
//
// Each frame slot owns one HostStage; it allocates once and is re-entrant across slots (no globals,
// unlike Triangle's file-scope state, triangle.cpp:541-550).
#pragma once
#include <cstdint>
#include <vector>

#include "common.cuh"

namespace elasb {

// Triangle-1.6-compatible divide-and-conquer Delaunay triangulator (alternating cuts) over integer
// points: same triangles, same corner order and same output order as the reference's
// triangulate("zQB") + writeelements (triangle.cpp:6160-6217, :7800-7853).
class Triangulator {
public:
    // x,y: n integer points (duplicates allowed, the first in sorted order survives,
    // triangle.cpp:6180-6195).  Appends (c1,c2,c3) index triples to `out` (cleared first).
    void run(const int32_t* x, const int32_t* y, int n, std::vector<int32_t>& out);

private:
    struct OTri { int t, o; };
    OTri sym(OTri a) const { int e = nbr_[3 * a.t + a.o]; return {e >> 2, e & 3}; }
    // (o + 1) % 3 and (o + 2) % 3 without a branch or a division: 2-bit fields of a constant
    static int plus1(int o) { return (0x09 >> (2 * o)) & 3; }     // 0 -> 1, 1 -> 2, 2 -> 0
    static int minus1(int o) { return (0x12 >> (2 * o)) & 3; }    // 0 -> 2, 1 -> 0, 2 -> 1
    static OTri lnext(OTri a) { return {a.t, plus1(a.o)}; }
    static OTri lprev(OTri a) { return {a.t, minus1(a.o)}; }
    int org(OTri a) const { return vtx_[3 * a.t + plus1(a.o)]; }
    int dest(OTri a) const { return vtx_[3 * a.t + minus1(a.o)]; }
    int apex(OTri a) const { return vtx_[3 * a.t + a.o]; }
    void set_org(OTri a, int v) { vtx_[3 * a.t + plus1(a.o)] = v; }
    void set_dest(OTri a, int v) { vtx_[3 * a.t + minus1(a.o)] = v; }
    void set_apex(OTri a, int v) { vtx_[3 * a.t + a.o] = v; }
    void bond(OTri a, OTri b) { nbr_[3 * a.t + a.o] = (b.t << 2) | b.o; nbr_[3 * b.t + b.o] = (a.t << 2) | a.o; }
    OTri make();
    int ccw(int a, int b, int c) const;
    int incircle(int a, int b, int c, int d) const;
    int random(unsigned choices);
    bool before(int a, int b, int axis) const;
    void sort(int* s, int n);
    void median(int* s, int n, int med, int axis);
    void alternate(int* s, int n, int axis);
    bool presorted_order(int n);
    void arrange(int* xs, int* ys, int n, int axis, int* out);
    void merge(OTri& farleft, OTri& innerleft, OTri& innerright, OTri& farright, int axis);
    void recurse(int* s, int n, int axis, OTri& farleft, OTri& farright);

    const int32_t* x_ = nullptr;
    const int32_t* y_ = nullptr;
    std::vector<int> nbr_, vtx_, order_, sx_, sy_, tmp_, count_;
    std::vector<uint8_t> side_;
    int ntri_ = 0;
    bool small_ = false;      // coordinate range below 2^14: incircle fits 64-bit integers
    uint64_t seed_ = 1;
};

struct HostStage {
    // outputs of run()
    int n_support = 0;
    std::vector<int32_t> support;            // (u,v,d) triples, elas.cpp:505-517 order
    std::vector<int32_t> tri[2];             // (c1,c2,c3) per image
    std::vector<float> planes[2];            // (t1a,t1b,t1c,t2a,t2b,t2c) per triangle
    std::vector<TriRaster> raster[2];
    // scan-conversion work units for k_raster: {triangle | image << 30, column chunk | row band << 16};
    // a big triangle becomes several units so that no single warp rasterises a large area alone
    std::vector<int32_t> units;
    std::vector<int16_t> dcan_incon;         // lattice after removeInconsistentSupportPoints (stage dump)

    // dcan: the candidate lattice [Hc][Wc] as produced by K2; filtered in place.
    // Returns the number of support points (callers stop at < 3, elas.cpp:69-75).
    // with_planes: also fit the disparity planes and raster records on the host (the frame path does
    // that on the device, k_planes; the host version serves elas_b200_host_stage and its CPU tests).
    int run(const FrameGeom& g, const elas_b200_params& p, int16_t* dcan, bool keep_stages, bool with_planes);

private:
    Triangulator delaunay_;
    std::vector<int32_t> px_, py_, col_fill_, cells_;
    std::vector<int16_t> pad_;
};

}  // namespace elasb
