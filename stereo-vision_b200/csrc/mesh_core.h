// Mesh stage, data-parallel formulation: the three in-place lattice filters, the support list, the
// Triangle-compatible Delaunay triangulation of both images and the scan-conversion work units, as
// PHASES that the threads of one CTA execute between barriers (k_mesh.cu).  A phase function does the
// share of thread `tid` of `nthr`; threads of one phase never depend on each other's results (beyond
// monotone flags), so running a phase for tid = 0..nthr-1 one after the other gives the same result --
// that is how tests/mesh_emulate.cpp checks these phases on the CPU against the sequential host stage
// (host_stage.cc), which in turn is pinned to the reference (tests/test_oracle.py).
//
// Reference: elas.cpp:174-279 (filters), :505-517 (support list), :534-600 + triangle.cpp:5446-6217
// (divide-and-conquer Delaunay with alternating cuts), :7800-7853 (element order).
//
// Why the sequential reference code can be cut into parallel phases:
//  * removeInconsistentSupportPoints (in place, u outer / v inner): a cell is invalidated iff fewer than
//    incon_min_support cells of its window are valid, similar and -- if they precede it in scan order --
//    not invalidated themselves.  That is a recursive definition along the scan order, so it has exactly
//    one solution; "invalidate every cell the rule justifies given the invalidations known so far" is a
//    monotone iteration that reaches it from the empty set, in any update order (incon_round()).
//  * removeRedundantSupportPoints: the vertical pass only looks along a lattice column, the horizontal
//    pass only along a row, so columns (rows) are independent sequential scans.
//  * Triangle's divide-and-conquer: the recursion tree is a pure function of n (split at n >> 1), a
//    subtree of m vertices allocates exactly 2m-2 triangles (2 per edge leaf, 4 per triangle leaf, 2 per
//    merge, nothing is freed before the ghosts are removed), so every node's slice of the triangle pool
//    is known in advance and all nodes of one depth can run at once; merges only touch their own subtree.
//  * vertexsort + alternateaxes are randomised, but without duplicate points their result is unique: the
//    k-d order is built with stable partitions of the two sorted id lists (prefix sums).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define MESH_FN __host__ __device__ __forceinline__
#else
#define MESH_FN inline
#endif

namespace elasb {
namespace mesh {

constexpr int kPadC = 8, kPadR = 8;        // lattice padding: >= incon_window_size and >= the redundancy reach (5)
// bits of a valid lattice value (disparities stay below 4096) while the inconsistency filter runs
constexpr int kValueMask = 0x0FFF;
constexpr int kRemoved = 0x4000;           // invalidated by the inconsistency filter
constexpr int kRedunDist = 5, kRedunThr = 1;   // elas.cpp:501-502

struct Lattice {
    int Wc, Hc, pitch, step;               // pitch = Wc + 2*kPadC (int16 elements)
    int16_t* P;                            // padded lattice [Hc + 2*kPadR][pitch], padding = -1
};

MESH_FN int lat_index(const Lattice& L, int uc, int vc) { return (vc + kPadR) * L.pitch + uc + kPadC; }
MESH_FN int lat_elems(int Wc, int Hc) { return (Wc + 2 * kPadC) * (Hc + 2 * kPadR); }
MESH_FN int iabs(int x) { return x < 0 ? -x : x; }

// ---- phase L0: padded copy of the candidate lattice (K2's output, [Hc][Wc]) ----------------------------
MESH_FN void lattice_load(const Lattice& L, const int16_t* dcan, int tid, int nthr)
{
    const int rows = L.Hc + 2 * kPadR, total = rows * L.pitch;
    for (int i = tid; i < total; i += nthr) {
        const int r = i / L.pitch, c = i - r * L.pitch;
        const int vc = r - kPadR, uc = c - kPadC;
        L.P[i] = (vc >= 0 && vc < L.Hc && uc >= 0 && uc < L.Wc) ? dcan[vc * L.Wc + uc] : (int16_t)-1;
    }
}

// removeInconsistentSupportPoints (elas.cpp:174-209) invalidates a cell when fewer than incon_min_support cells
// of its window are valid, similar (|d - d'| <= incon_threshold) and -- if they precede it in scan order (u outer,
// v inner) -- not invalidated themselves.  With count0 = the supporters in the UNFILTERED lattice (the cell
// itself included) that reads: invalid(c) <=> count0(c) - #{invalidated earlier similar cells of the window} < need.
// Solved by propagation: cells with count0 < need seed a work list; every listed cell is invalidated and takes one
// supporter away from each LATER similar valid cell of its window; a cell whose count thereby drops below need
// joins the next list.  Counts only fall, a cell crosses the threshold once, the result is the unique solution of
// the recursive definition -- whatever the processing order.

// ---- phase C (its own kernel on the device, any number of CTAs): count0 of one cell of the unfiltered lattice --
MESH_FN int incon_count0(const int16_t* dcan, int Wc, int Hc, int uc, int vc, int win, int thr)
{
    const int x = dcan[vc * Wc + uc];
    if (x < 0) return 0;
    int support = 0;
    for (int v2 = (vc - win > 0 ? vc - win : 0); v2 <= vc + win && v2 < Hc; v2++)
        for (int u2 = (uc - win > 0 ? uc - win : 0); u2 <= uc + win && u2 < Wc; u2++) {
            const int y = dcan[v2 * Wc + u2];
            support += y >= 0 && iabs(x - y) <= thr;
        }
    return support;
}

// ---- phase L1a: the cells that fail on their own seed the first work list (entries = padded lattice indices) ---
template <class AtomicAdd>
MESH_FN void incon_seed(const Lattice& L, const int32_t* cnt, int need, int32_t* list, int* list_n, int tid, int nthr, AtomicAdd atomic_add)
{
    const int cells = L.Wc * L.Hc;
    for (int i = tid; i < cells; i += nthr) {
        const int vc = i / L.Wc, uc = i - vc * L.Wc;
        const int at = lat_index(L, uc, vc);
        if (L.P[at] >= 0 && cnt[i] < need) list[atomic_add(list_n, 1)] = at;
    }
}
// ---- phase L1b (repeated until the next list stays empty): invalidate the listed cells, propagate --------------
// atomic_add(int*, int) returns the previous value; or16(int16_t*, bits) sets bits atomically.
template <class AtomicAdd, class Or16>
MESH_FN void incon_propagate(const Lattice& L, int32_t* cnt, int win, int thr, int need, const int32_t* cur, int n_cur,
                             int32_t* next, int* n_next, int tid, int nthr, AtomicAdd atomic_add, Or16 or16)
{
    for (int k = tid; k < n_cur; k += nthr) {
        const int at = cur[k];
        int16_t* c = L.P + at;
        const int x = *c & kValueMask;
        or16(c, kRemoved);
        const int vc = at / L.pitch - kPadR, uc = at - (vc + kPadR) * L.pitch - kPadC;
        for (int dv = -win; dv <= win; dv++)
            for (int du = (dv > 0 ? 0 : 1); du <= win; du++) {              // later in scan order only
                const int y = c[dv * L.pitch + du];
                if (y < 0 || iabs(x - (y & kValueMask)) > thr) continue;      // padding, invalid or not similar
                if (atomic_add(cnt + (vc + dv) * L.Wc + (uc + du), -1) == need) next[atomic_add(n_next, 1)] = at + dv * L.pitch + du;
            }
    }
}

// ---- phase L2: invalidated cells become -1; optional dump of the lattice after this filter -----------------
MESH_FN void incon_finish(const Lattice& L, int16_t* dump, int tid, int nthr)
{
    const int cells = L.Wc * L.Hc;
    for (int i = tid; i < cells; i += nthr) {
        const int vc = i / L.Wc, uc = i - vc * L.Wc;
        int16_t* c = L.P + lat_index(L, uc, vc);
        if (*c >= 0 && (*c & kRemoved)) *c = -1;
        if (dump) dump[i] = *c;
    }
}

// ---- phases L3 / L4: removeRedundantSupportPoints, elas.cpp:213-279 -------------------------------------
// vertical = true: thread per lattice column scanning v; false: thread per lattice row scanning u.  In both
// cases the cells behind the scan position already hold this pass's result, the cells ahead its input.
MESH_FN void redundant_pass(const Lattice& L, bool vertical, int tid, int nthr)
{
    const int lines = vertical ? L.Wc : L.Hc, len = vertical ? L.Hc : L.Wc;
    const int stride = vertical ? L.pitch : 1;
    for (int line = tid; line < lines; line += nthr) {
        int16_t* c = L.P + (vertical ? lat_index(L, line, 0) : lat_index(L, 0, line));
        for (int k = 0; k < len; k++, c += stride) {
            const int d = *c;
            if (d < 0) continue;
            bool back = false, fwd = false;
            for (int j = 1; j <= kRedunDist; j++) { const int x = c[-j * stride]; back |= x >= 0 && iabs(d - x) <= kRedunThr; }
            if (!back) continue;
            for (int j = 1; j <= kRedunDist; j++) { const int x = c[j * stride]; fwd |= x >= 0 && iabs(d - x) <= kRedunThr; }
            if (fwd) *c = -1;
        }
    }
}

// ---- phase L5a: survivors per lattice column (lattice row 0 / column 0 excluded, elas.cpp:506-508) ------
MESH_FN void support_count(const Lattice& L, int32_t* col_count, int tid, int nthr)
{
    for (int uc = tid; uc < L.Wc; uc += nthr) {
        int n = 0;
        if (uc >= 1)
            for (int vc = 1; vc < L.Hc; vc++) n += L.P[lat_index(L, uc, vc)] >= 0;
        col_count[uc] = n;
    }
}
// ---- phase L5b: the support list in u-outer / v-inner order (elas.cpp:505-517); col_off = exclusive scan --
MESH_FN void support_write(const Lattice& L, const int32_t* col_off, int32_t* support, int tid, int nthr)
{
    for (int uc = 1 + tid; uc < L.Wc; uc += nthr) {
        int at = col_off[uc];
        for (int vc = 1; vc < L.Hc; vc++) {
            const int d = L.P[lat_index(L, uc, vc)];
            if (d < 0) continue;
            support[3 * at] = uc * L.step; support[3 * at + 1] = vc * L.step; support[3 * at + 2] = d;
            at++;
        }
    }
}
// ---- phase L6: the filtered lattice back to [Hc][Wc] ----------------------------------------------------
MESH_FN void lattice_store(const Lattice& L, int16_t* dcan, int tid, int nthr)
{
    const int cells = L.Wc * L.Hc;
    for (int i = tid; i < cells; i += nthr) {
        const int vc = i / L.Wc, uc = i - vc * L.Wc;
        dcan[i] = L.P[lat_index(L, uc, vc)];
    }
}

// =============================================================================================
// Delaunay
// =============================================================================================
// An oriented triangle (edge org->dest of triangle t, apex opposite) is ONE int e = 4 t + o, o in {0,1,2}: it is
// at the same time the index of its neighbour link and of its apex in the per-triangle records of four ints
// (slot 3 unused) -- the merges are chains of dependent loads executed by single threads, every instruction
// of address arithmetic saved there is run time.
typedef int OTri;

struct Mesh {
    int n;                                  // vertices
    const uint32_t* xy;                     // coordinates by vertex id, x << 16 | y (both below 2^14)
    int32_t* s;                             // vertex ids in the alternating-cut order (triangle.cpp:6197-6206)
    int32_t* nbr; int32_t* vtx;             // 4 ints per triangle (3 links / 3 vertices), 2n-2 triangles; vertex -1 = ghost
    int32_t* hull;                          // [2n]: (farleft, farright) handles of the subtree that starts at s[lo]
};

MESH_FN int plus1(int o) { return (0x09 >> (2 * o)) & 3; }      // 0 -> 1, 1 -> 2, 2 -> 0
MESH_FN int minus1(int o) { return (0x12 >> (2 * o)) & 3; }     // 0 -> 2, 1 -> 0, 2 -> 1
MESH_FN OTri lnext(OTri a) { return (a & ~3) | plus1(a & 3); }
MESH_FN OTri lprev(OTri a) { return (a & ~3) | minus1(a & 3); }
MESH_FN int enc(OTri a) { return a; }
MESH_FN OTri dec(int e) { return e; }
MESH_FN OTri sym(const Mesh& m, OTri a) { return m.nbr[a]; }
MESH_FN int org(const Mesh& m, OTri a) { return m.vtx[lnext(a)]; }
MESH_FN int dest(const Mesh& m, OTri a) { return m.vtx[lprev(a)]; }
MESH_FN int apex(const Mesh& m, OTri a) { return m.vtx[a]; }
MESH_FN void set_org(const Mesh& m, OTri a, int v) { m.vtx[lnext(a)] = v; }
MESH_FN void set_dest(const Mesh& m, OTri a, int v) { m.vtx[lprev(a)] = v; }
MESH_FN void set_apex(const Mesh& m, OTri a, int v) { m.vtx[a] = v; }
MESH_FN void bond(const Mesh& m, OTri a, OTri b) { m.nbr[a] = b; m.nbr[b] = a; }
MESH_FN OTri make(const Mesh& m, int t)
{
    for (int i = 0; i < 3; i++) { m.nbr[4 * t + i] = -1; m.vtx[4 * t + i] = -1; }
    return 4 * t;
}

// A vertex with its coordinates in registers: the merge loop keeps the four corners of the knitting edge and
// the candidate apex this way, so a predicate costs one load (the new vertex) instead of eight.
struct PV { int id, x, y; };
MESH_FN PV pv(const Mesh& m, int id)
{
    const uint32_t c = m.xy[id];
    return {id, (int)(c >> 16), (int)(c & 0xFFFFu)};
}
// exact predicates (triangle.cpp:2706, :3334 return exact signs): coordinates are integers below 2^14, so every
// difference is below 2^14, every product of two below 2^28 (32-bit), the in-circle determinant below 2^60
MESH_FN int ccw(const PV& a, const PV& b, const PV& c)
{
    const int l = (a.x - c.x) * (b.y - c.y), r = (a.y - c.y) * (b.x - c.x);
    return (l > r) - (l < r);
}
MESH_FN int incircle(const PV& a, const PV& b, const PV& c, const PV& d)
{
    const int adx = a.x - d.x, ady = a.y - d.y, bdx = b.x - d.x, bdy = b.y - d.y, cdx = c.x - d.x, cdy = c.y - d.y;
    const long long det = (long long)(adx * adx + ady * ady) * (bdx * cdy - cdx * bdy) +
                          (long long)(bdx * bdx + bdy * bdy) * (cdx * ady - adx * cdy) +
                          (long long)(cdx * cdx + cdy * cdy) * (adx * bdy - bdx * ady);
    return (det > 0) - (det < 0);
}

// ---- alternating-cut order (vertexsort + alternateaxes without duplicates) ----------------------------
// xs / ys: the vertex ids sorted by (x,y) / (y,x); posx / posy: their inverse permutations; seg_lo / seg_n:
// the segment [lo, lo+n) of the current level a vertex lies in (both lists hold the same set per segment).
// One level: segments longer than 3 are cut at n >> 1 along the level's axis (0: by x, 1: by y); the list
// sorted along the other axis is stably partitioned.  side[] is indexed by POSITION in the other list.
struct Order {
    int n;
    int32_t* xs; int32_t* ys; int32_t* posx; int32_t* posy; int32_t* seg_lo; int32_t* seg_n;
    int32_t* side;       // [n]   1 = goes to the upper part, by position in the level's "other" list
    int32_t* scan;       // [n]   inclusive prefix sum of side[]
    int32_t* tmp;        // [n]   the partitioned other list before it is copied back
};

MESH_FN void order_init(const Order& o, int tid, int nthr)
{
    for (int i = tid; i < o.n; i += nthr) {
        o.posx[o.xs[i]] = i; o.posy[o.ys[i]] = i;
        o.seg_lo[i] = 0; o.seg_n[i] = o.n;
    }
}
// phase A of a level: side flags; returns true if this thread saw a segment that is still being cut
MESH_FN bool order_flags(const Order& o, int axis, int tid, int nthr)
{
    bool cutting = false;
    const int32_t* other = axis ? o.xs : o.ys;
    const int32_t* pos_primary = axis ? o.posy : o.posx;
    for (int i = tid; i < o.n; i += nthr) {
        const int id = other[i];
        const int lo = o.seg_lo[id], n = o.seg_n[id];
        int f = 0;
        if (n > 3) { f = pos_primary[id] - lo >= (n >> 1); cutting = true; }
        o.side[i] = f;
    }
    return cutting;
}
// phase B (after the inclusive scan of side[] into scan[]): new positions in the other list
MESH_FN void order_scatter(const Order& o, int axis, int tid, int nthr)
{
    const int32_t* other = axis ? o.xs : o.ys;
    for (int i = tid; i < o.n; i += nthr) {
        const int id = other[i];
        const int lo = o.seg_lo[id], n = o.seg_n[id];
        int to = i;
        if (n > 3) {
            const int divider = n >> 1;
            const int ones_before = o.scan[i] - o.side[i] - (lo > 0 ? o.scan[lo - 1] : 0);
            to = o.side[i] ? lo + divider + ones_before : lo + (i - lo) - ones_before;
        }
        o.tmp[to] = id;
    }
}
// phase C: copy back, new inverse permutation, descend into the sub-segment
MESH_FN void order_commit(const Order& o, int axis, int tid, int nthr)
{
    int32_t* other = axis ? o.xs : o.ys;
    int32_t* pos_other = axis ? o.posx : o.posy;
    const int32_t* pos_primary = axis ? o.posy : o.posx;
    for (int i = tid; i < o.n; i += nthr) {
        const int id = o.tmp[i];
        other[i] = id; pos_other[id] = i;
        const int lo = o.seg_lo[id], n = o.seg_n[id];
        if (n > 3) {
            const int divider = n >> 1;
            if (pos_primary[id] - lo >= divider) { o.seg_lo[id] = lo + divider; o.seg_n[id] = n - divider; }
            else o.seg_n[id] = divider;
        }
    }
}

// ---- divide and conquer by depth ------------------------------------------------------------------------
struct Node { int lo, n, pool, axis; bool exists; };
// node j (0 <= j < 2^depth) of the recursion tree over s[0..n): divconqrecurse splits at n >> 1 and
// alternates the axis starting with 0 (triangle.cpp:6213); pool = first triangle of the subtree's slice
MESH_FN Node node_at(int n, int depth, int j)
{
    Node nd = {0, n, 0, 0, true};
    for (int level = depth - 1; level >= 0; level--) {
        if (nd.n <= 3) { nd.exists = false; return nd; }                  // an ancestor is a leaf
        const int divider = nd.n >> 1;
        if ((j >> level) & 1) { nd.lo += divider; nd.pool += 2 * divider - 2; nd.n -= divider; }
        else nd.n = divider;
        nd.axis ^= 1;
    }
    return nd;
}
MESH_FN int tree_depth(int n) { int d = 0; while (n > 3) { n = (n + 1) >> 1; d++; } return d; }    // deepest level (largest child)

// leaves: two vertices = an edge with two ghost triangles (triangle.cpp:5965-5991), three = one triangle and
// three ghosts, or two edges when collinear (:5992-6088)
MESH_FN void build_leaf(const Mesh& m, const Node& nd, OTri& farleft, OTri& farright)
{
    const int32_t* s = m.s + nd.lo;
    int t = nd.pool;
    if (nd.n == 2) {
        farleft = make(m, t);
        set_org(m, farleft, s[0]); set_dest(m, farleft, s[1]);
        farright = make(m, t + 1);
        set_org(m, farright, s[1]); set_dest(m, farright, s[0]);
        bond(m, farleft, farright);
        farleft = lprev(farleft); farright = lnext(farright);
        bond(m, farleft, farright);
        farleft = lprev(farleft); farright = lnext(farright);
        bond(m, farleft, farright);
        farleft = lprev(farright);
        return;
    }
    OTri mid = make(m, t), t1 = make(m, t + 1), t2 = make(m, t + 2), t3 = make(m, t + 3);
    const int area = ccw(pv(m, s[0]), pv(m, s[1]), pv(m, s[2]));
    if (area == 0) {
        set_org(m, mid, s[0]); set_dest(m, mid, s[1]);
        set_org(m, t1, s[1]);  set_dest(m, t1, s[0]);
        set_org(m, t2, s[2]);  set_dest(m, t2, s[1]);
        set_org(m, t3, s[1]);  set_dest(m, t3, s[2]);
        bond(m, mid, t1); bond(m, t2, t3);
        mid = lnext(mid); t1 = lprev(t1); t2 = lnext(t2); t3 = lprev(t3);
        bond(m, mid, t3); bond(m, t1, t2);
        mid = lnext(mid); t1 = lprev(t1); t2 = lnext(t2); t3 = lprev(t3);
        bond(m, mid, t1); bond(m, t2, t3);
        farleft = t1;
        farright = t2;
    } else {
        const int p = area > 0 ? s[1] : s[2], q = area > 0 ? s[2] : s[1];
        set_org(m, mid, s[0]); set_dest(m, t1, s[0]); set_org(m, t3, s[0]);
        set_dest(m, mid, p);   set_org(m, t1, p);     set_dest(m, t2, p);
        set_apex(m, mid, q);   set_org(m, t2, q);     set_dest(m, t3, q);
        bond(m, mid, t1);
        mid = lnext(mid);
        bond(m, mid, t2);
        mid = lnext(mid);
        bond(m, mid, t3);
        t1 = lprev(t1); t2 = lnext(t2);
        bond(m, t1, t2);
        t1 = lprev(t1); t3 = lprev(t3);
        bond(m, t1, t3);
        t2 = lnext(t2); t3 = lprev(t3);
        bond(m, t2, t3);
        farleft = t1;
        farright = area > 0 ? t2 : lnext(farleft);
    }
}

// mergehulls, triangle.cpp:5638-5934: knits the hulls of two adjacent subtrees.  base_t / top_t = the two
// triangles the merge allocates (first and last).  All tie-breaks are the strict comparisons of the
// reference; on co-circular quads the LEFT candidate wins (:5908-5910).
MESH_FN void merge_hulls(const Mesh& m, OTri& farleft, OTri innerleft, OTri innerright, OTri& farright, int axis,
                         int base_t, int top_t)
{
    PV ild = pv(m, dest(m, innerleft)), ila = pv(m, apex(m, innerleft));
    PV iro = pv(m, org(m, innerright)), ira = pv(m, apex(m, innerright));

    if (axis == 1) {
        // horizontal cut: re-aim the four hull handles at the bottom-/top-most vertices (:5666-5704)
        PV flp = pv(m, org(m, farleft)), fla = pv(m, apex(m, farleft));
        PV frp = pv(m, dest(m, farright));
        while (fla.y < flp.y) {
            farleft = sym(m, lnext(farleft));
            flp = fla;
            fla = pv(m, apex(m, farleft));
        }
        OTri check = sym(m, innerleft);
        PV cv = pv(m, apex(m, check));
        while (cv.y > ild.y) {
            innerleft = lnext(check);
            ila = ild;
            ild = cv;
            check = sym(m, innerleft);
            cv = pv(m, apex(m, check));
        }
        while (ira.y < iro.y) {
            innerright = sym(m, lnext(innerright));
            iro = ira;
            ira = pv(m, apex(m, innerright));
        }
        check = sym(m, farright);
        cv = pv(m, apex(m, check));
        while (cv.y > frp.y) {
            farright = lnext(check);
            frp = cv;
            check = sym(m, farright);
            cv = pv(m, apex(m, check));
        }
    }

    // lower common tangent (:5706-5726)
    for (bool changed = true; changed;) {
        changed = false;
        if (ccw(ild, ila, iro) > 0) {
            innerleft = sym(m, lprev(innerleft));
            ild = ila;
            ila = pv(m, apex(m, innerleft));
            changed = true;
        }
        if (ccw(ira, iro, ild) > 0) {
            innerright = sym(m, lnext(innerright));
            iro = ira;
            ira = pv(m, apex(m, innerright));
            changed = true;
        }
    }

    OTri leftcand = sym(m, innerleft), rightcand = sym(m, innerright);
    OTri base = make(m, base_t);                         // bottom bounding triangle (:5731-5738)
    bond(m, base, innerleft);
    base = lnext(base);
    bond(m, base, innerright);
    base = lnext(base);
    set_org(m, base, iro.id);
    set_dest(m, base, ild.id);
    if (ild.id == org(m, farleft)) farleft = lnext(base);   // :5745-5752
    if (iro.id == dest(m, farright)) farright = lprev(base);

    PV lowerleft = ild, lowerright = iro;
    PV upperleft = pv(m, apex(m, leftcand)), upperright = pv(m, apex(m, rightcand));

    for (;;) {
        const bool leftdone = ccw(upperleft, lowerleft, lowerright) <= 0;     // :5765-5768
        const bool rightdone = ccw(upperright, lowerleft, lowerright) <= 0;
        if (leftdone && rightdone) {
            OTri top = make(m, top_t);                   // top bounding triangle (:5771-5780)
            set_org(m, top, lowerleft.id);
            set_dest(m, top, lowerright.id);
            bond(m, top, base);
            top = lnext(top);
            bond(m, top, rightcand);
            top = lnext(top);
            bond(m, top, leftcand);
            if (axis == 1) {
                // restore the handles to the left-/right-most vertices (:5786-5809)
                PV flp = pv(m, org(m, farleft));
                PV frp = pv(m, dest(m, farright)), fra = pv(m, apex(m, farright));
                OTri check = sym(m, farleft);
                PV cv = pv(m, apex(m, check));
                while (cv.x < flp.x) {
                    farleft = lprev(check);
                    flp = cv;
                    check = sym(m, farleft);
                    cv = pv(m, apex(m, check));
                }
                while (fra.x > frp.x) {
                    farright = sym(m, lprev(farright));
                    frp = fra;
                    fra = pv(m, apex(m, farright));
                }
            }
            return;
        }
        if (!leftdone) {
            // flip away left-hull edges that are not Delaunay w.r.t. the knitting edge (:5813-5859)
            OTri next = sym(m, lprev(leftcand));
            int nextapex = apex(m, next);
            while (nextapex >= 0) {
                const PV na = pv(m, nextapex);
                if (!(incircle(lowerleft, lowerright, upperleft, na) > 0)) break;
                next = lnext(next);
                const OTri topcasing = sym(m, next);
                next = lnext(next);
                const OTri sidecasing = sym(m, next);
                bond(m, next, topcasing);
                bond(m, leftcand, sidecasing);
                leftcand = lnext(leftcand);
                const OTri outercasing = sym(m, leftcand);
                next = lprev(next);
                bond(m, next, outercasing);
                set_org(m, leftcand, lowerleft.id);
                set_dest(m, leftcand, -1);
                set_apex(m, leftcand, nextapex);
                set_org(m, next, -1);
                set_dest(m, next, upperleft.id);
                set_apex(m, next, nextapex);
                upperleft = na;
                next = sidecasing;
                nextapex = apex(m, next);
            }
        }
        if (!rightdone) {
            // same on the right hull (:5861-5907)
            OTri next = sym(m, lnext(rightcand));
            int nextapex = apex(m, next);
            while (nextapex >= 0) {
                const PV na = pv(m, nextapex);
                if (!(incircle(lowerleft, lowerright, upperright, na) > 0)) break;
                next = lprev(next);
                const OTri topcasing = sym(m, next);
                next = lprev(next);
                const OTri sidecasing = sym(m, next);
                bond(m, next, topcasing);
                bond(m, rightcand, sidecasing);
                rightcand = lprev(rightcand);
                const OTri outercasing = sym(m, rightcand);
                next = lnext(next);
                bond(m, next, outercasing);
                set_org(m, rightcand, -1);
                set_dest(m, rightcand, lowerright.id);
                set_apex(m, rightcand, nextapex);
                set_org(m, next, upperright.id);
                set_dest(m, next, -1);
                set_apex(m, next, nextapex);
                upperright = na;
                next = sidecasing;
                nextapex = apex(m, next);
            }
        }
        // choose the next tooth; on co-circular quads the LEFT candidate wins (:5908-5910)
        if (leftdone || (!rightdone && incircle(upperleft, lowerleft, lowerright, upperright) > 0)) {
            bond(m, base, rightcand);
            base = lprev(rightcand);
            set_dest(m, base, lowerleft.id);
            lowerright = upperright;
            rightcand = sym(m, base);
            upperright = pv(m, apex(m, rightcand));
        } else {
            bond(m, base, leftcand);
            base = lnext(leftcand);
            set_org(m, base, lowerright.id);
            lowerleft = upperleft;
            leftcand = sym(m, base);
            upperleft = pv(m, apex(m, leftcand));
        }
    }
}

// ---- phase D(depth): every node of one depth, deepest level first (divconqrecurse, :5953-6103) ---------
MESH_FN void triangulate_depth(const Mesh& m, int depth, int tid, int nthr)
{
    for (int j = tid; j < (1 << depth); j += nthr) {
        const Node nd = node_at(m.n, depth, j);
        if (!nd.exists) continue;
        OTri farleft, farright;
        if (nd.n <= 3) build_leaf(m, nd, farleft, farright);
        else {
            const int divider = nd.n >> 1;
            farleft = dec(m.hull[2 * nd.lo]);
            const OTri innerleft = dec(m.hull[2 * nd.lo + 1]);
            const OTri innerright = dec(m.hull[2 * (nd.lo + divider)]);
            farright = dec(m.hull[2 * (nd.lo + divider) + 1]);
            merge_hulls(m, farleft, innerleft, innerright, farright, nd.axis, nd.pool + 2 * nd.n - 4, nd.pool + 2 * nd.n - 3);
        }
        m.hull[2 * nd.lo] = enc(farleft);
        m.hull[2 * nd.lo + 1] = enc(farright);
    }
}

// ---- output: removeghosts frees exactly the triangles that hold the ghost vertex (:6105-6148); writeelements
// walks the pool in allocation order, corners (org, dest, apex) at orientation 0 (:7834-7853) --------------
MESH_FN void real_flags(const Mesh& m, int32_t* flag, int tid, int nthr)
{
    const int nt = 2 * m.n - 2;
    for (int t = tid; t < nt; t += nthr)
        flag[t] = m.vtx[4 * t] >= 0 && m.vtx[4 * t + 1] >= 0 && m.vtx[4 * t + 2] >= 0;
}
// scan = inclusive prefix sum of flag[]; tri_out = (c1,c2,c3) triples
MESH_FN void write_triangles(const Mesh& m, const int32_t* flag, const int32_t* scan, int32_t* tri_out, int tid, int nthr)
{
    const int nt = 2 * m.n - 2;
    for (int t = tid; t < nt; t += nthr) {
        if (!flag[t]) continue;
        int32_t* o = tri_out + 3 * (scan[t] - 1);
        o[0] = m.vtx[4 * t + 1]; o[1] = m.vtx[4 * t + 2]; o[2] = m.vtx[4 * t];
    }
}

// ---- scan-conversion work units: 32-column chunks x band_rows-row bands of each triangle's bounding box ----
struct Box { int u_lo, v_lo, chunks, bands; };
MESH_FN Box raster_box(const int32_t* x, const int32_t* y, const int32_t* tri, int W, int H, int band_rows)
{
    const int a = tri[0], b = tri[1], c = tri[2];
    const int xmin = x[a] < x[b] ? (x[a] < x[c] ? x[a] : x[c]) : (x[b] < x[c] ? x[b] : x[c]);
    const int xmax = x[a] > x[b] ? (x[a] > x[c] ? x[a] : x[c]) : (x[b] > x[c] ? x[b] : x[c]);
    const int ymin = y[a] < y[b] ? (y[a] < y[c] ? y[a] : y[c]) : (y[b] < y[c] ? y[b] : y[c]);
    const int ymax = y[a] > y[b] ? (y[a] > y[c] ? y[a] : y[c]) : (y[b] > y[c] ? y[b] : y[c]);
    Box bx;
    bx.u_lo = xmin > 0 ? xmin : 0;
    const int u_hi = xmax < W ? xmax : W;                                 // columns [u_lo, u_hi)
    // one row of slack below the smallest corner row: an edge line evaluated in float may truncate to it
    bx.v_lo = ymin - 1 > 0 ? ymin - 1 : 0;
    const int v_hi = ymax + 1 < H ? ymax + 1 : H;                         // rows [v_lo, v_hi)
    bx.chunks = (u_hi - bx.u_lo + 31) / 32;
    bx.bands = (v_hi - bx.v_lo + band_rows - 1) / band_rows;
    if (bx.chunks <= 0 || bx.bands <= 0) bx.chunks = bx.bands = 0;
    return bx;
}
MESH_FN void unit_counts(const int32_t* x, const int32_t* y, const int32_t* tri, int nt, int W, int H, int band_rows,
                         int32_t* count, int tid, int nthr)
{
    for (int t = tid; t < nt; t += nthr) {
        const Box bx = raster_box(x, y, tri + 3 * t, W, H, band_rows);
        count[t] = bx.chunks * bx.bands;
    }
}
// scan = inclusive prefix sum of count[]; units = {triangle | image << 30, chunk | band << 16}.  A triangle whose
// units do not fit below unit_cap is listed in overflow[] instead (scan-converted by one warp on its own);
// ovf_scan = inclusive prefix sum of the overflow flags.
MESH_FN bool unit_overflows(const int32_t* scan, int t, int unit_cap) { return scan[t] > unit_cap; }
MESH_FN void write_units(const int32_t* x, const int32_t* y, const int32_t* tri, int nt, int W, int H, int band_rows,
                         int image, const int32_t* count, const int32_t* scan, int unit_cap, int32_t* units,
                         int tid, int nthr)
{
    for (int t = tid; t < nt; t += nthr) {
        if (count[t] == 0 || unit_overflows(scan, t, unit_cap)) continue;
        const Box bx = raster_box(x, y, tri + 3 * t, W, H, band_rows);
        int32_t* o = units + 2 * (scan[t] - count[t]);
        for (int ch = 0; ch < bx.chunks; ch++)
            for (int bd = 0; bd < bx.bands; bd++) { *o++ = t | (image << 30); *o++ = ch | (bd << 16); }
    }
}

}  // namespace mesh
}  // namespace elasb
