// Host middle stage: lattice filters, support list, Triangle-compatible Delaunay, disparity planes,
// raster records.  See host_stage.h.  Compiled with -ffp-contract=off: the double/float expressions
// must round exactly as the reference's x86-64 SSE code does.
#include "host_stage.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include <emmintrin.h>

namespace elasb {

// =============================================================================================
// Triangulator: divide-and-conquer Delaunay with alternating cuts, Triangle-compatible.
//
// Structure (triangle.cpp:5638-6217): triangles live in a pool in allocation order; each has three
// neighbour links (triangle index * 4 + orientation) and three vertices, vertex -1 being the ghost
// vertex of the "bounding" triangles that surround the hull.  An oriented triangle (t, o) names the
// edge org->dest with apex opposite.  Merging two hulls converts bounding triangles into real ones
// and vice versa in place, which is what fixes the output ORDER; tie-breaks are the strict
// comparisons of mergehulls.  The predicates are exact integer determinants (the reference's
// adaptive float predicates return exact signs; ELAS coordinates are small integers).
// =============================================================================================

Triangulator::OTri Triangulator::make()
{
    const int t = ntri_++;
    for (int i = 0; i < 3; i++) { nbr_[3 * t + i] = -1; vtx_[3 * t + i] = -1; }
    return {t, 0};
}

int Triangulator::ccw(int a, int b, int c) const
{
    const int64_t l = (int64_t)(x_[a] - x_[c]) * (y_[b] - y_[c]);
    const int64_t r = (int64_t)(y_[a] - y_[c]) * (x_[b] - x_[c]);
    return (l > r) - (l < r);
}

int Triangulator::incircle(int a, int b, int c, int d) const
{
    const int64_t adx = x_[a] - x_[d], ady = y_[a] - y_[d];
    const int64_t bdx = x_[b] - x_[d], bdy = y_[b] - y_[d];
    const int64_t cdx = x_[c] - x_[d], cdy = y_[c] - y_[d];
    if (small_) {
        // coordinate differences below 2^14: every product below 2^58, the sum below 2^60
        const int64_t det = (adx * adx + ady * ady) * (bdx * cdy - cdx * bdy) +
                            (bdx * bdx + bdy * bdy) * (cdx * ady - adx * cdy) +
                            (cdx * cdx + cdy * cdy) * (adx * bdy - bdx * ady);
        return (det > 0) - (det < 0);
    }
    const __int128 det = (__int128)(adx * adx + ady * ady) * (bdx * cdy - cdx * bdy) +
                         (__int128)(bdx * bdx + bdy * bdy) * (cdx * ady - adx * cdy) +
                         (__int128)(cdx * cdx + cdy * cdy) * (adx * bdy - bdx * ady);
    return (det > 0) - (det < 0);
}

int Triangulator::random(unsigned choices)              // triangle.cpp:4045-4049
{
    seed_ = (seed_ * 1366u + 150889u) % 714025u;
    return (int)(seed_ / (714025u / choices + 1));
}

bool Triangulator::before(int a, int b, int axis) const
{
    const int32_t a1 = axis ? y_[a] : x_[a], a2 = axis ? x_[a] : y_[a];
    const int32_t b1 = axis ? y_[b] : x_[b], b2 = axis ? x_[b] : y_[b];
    return a1 < b1 || (a1 == b1 && a2 < b2);
}

void Triangulator::sort(int* s, int n)                   // vertexsort, triangle.cpp:5446-5499
{
    if (n == 2) {
        if (before(s[1], s[0], 0)) std::swap(s[0], s[1]);
        return;
    }
    const int pivot = s[random((unsigned)n)];
    int left = -1, right = n;
    while (left < right) {
        do { left++; } while (left <= right && before(s[left], pivot, 0));
        do { right--; } while (left <= right && before(pivot, s[right], 0));
        if (left < right) std::swap(s[left], s[right]);
    }
    if (left > 1) sort(s, left);
    if (right < n - 2) sort(s + right + 1, n - right - 1);
}

void Triangulator::median(int* s, int n, int med, int axis)   // vertexmedian, :5513-5569
{
    if (n == 2) {
        if (before(s[1], s[0], axis)) std::swap(s[0], s[1]);
        return;
    }
    const int pivot = s[random((unsigned)n)];
    int left = -1, right = n;
    while (left < right) {
        do { left++; } while (left <= right && before(s[left], pivot, axis));
        do { right--; } while (left <= right && before(pivot, s[right], axis));
        if (left < right) std::swap(s[left], s[right]);
    }
    if (left > med) median(s, left, med, axis);
    if (right < med - 1) median(s + right + 1, n - right - 1, med - right - 1, axis);
}

void Triangulator::alternate(int* s, int n, int axis)     // alternateaxes, :5582-5601
{
    const int divider = n >> 1;
    if (n <= 3) axis = 0;
    median(s, n, divider, axis);
    if (n - divider >= 2) {
        if (divider >= 2) alternate(s, divider, 1 - axis);
        alternate(s + divider, n - divider, 1 - axis);
    }
}

// Fast path for the vertex order.  vertexsort + alternateaxes are randomised, but without duplicate
// points every comparison is between distinct keys, so the array they produce is unique: the set
// split at each level is fixed by the keys and the leaves (<= 3 vertices) end up sorted by x.  That
// arrangement is built here from two counting-sorted id lists (by (x,y) and by (y,x)) with stable
// partitions, k-d-tree style: O(n log n) sequential passes instead of ~55 unpredictable indirect
// comparisons per vertex.  Returns false (caller falls back to the literal quicksort/quickselect,
// whose pivot sequence decides which duplicate survives) when two points coincide.
bool Triangulator::presorted_order(int n)
{
    int32_t xmin = x_[0], xmax = x_[0], ymin = y_[0], ymax = y_[0];
    for (int i = 1; i < n; i++) {
        xmin = std::min(xmin, x_[i]); xmax = std::max(xmax, x_[i]);
        ymin = std::min(ymin, y_[i]); ymax = std::max(ymax, y_[i]);
    }
    const int64_t rx = (int64_t)xmax - xmin + 1, ry = (int64_t)ymax - ymin + 1;
    if (rx > (1 << 22) || ry > (1 << 22)) return false;
    sx_.resize(n); sy_.resize(n); tmp_.resize(n); side_.resize(n);
    // stable counting sort of ids 0..n-1 by key
    auto counting = [&](const int* in, int* out, const int32_t* key, int32_t kmin, int64_t range) {
        count_.assign((size_t)range + 1, 0);
        for (int i = 0; i < n; i++) count_[key[in[i]] - kmin + 1]++;
        for (int64_t k = 0; k < range; k++) count_[k + 1] += count_[k];
        for (int i = 0; i < n; i++) out[count_[key[in[i]] - kmin]++] = in[i];
    };
    int* ids = order_.data();
    for (int i = 0; i < n; i++) ids[i] = i;
    counting(ids, tmp_.data(), y_, ymin, ry);            // by y, then stably by x  => (x, y)
    counting(tmp_.data(), sx_.data(), x_, xmin, rx);
    counting(ids, tmp_.data(), x_, xmin, rx);            // by x, then stably by y  => (y, x)
    counting(tmp_.data(), sy_.data(), y_, ymin, ry);
    for (int i = 1; i < n; i++)
        if (x_[sx_[i]] == x_[sx_[i - 1]] && y_[sx_[i]] == y_[sx_[i - 1]]) return false;
    arrange(sx_.data(), sy_.data(), n, 0, ids);
    return true;
}

// xs / ys: the same n ids sorted by (x,y) / (y,x).  Writes the alternating-cut order to out[0..n).
void Triangulator::arrange(int* xs, int* ys, int n, int axis, int* out)
{
    if (n <= 3) {                                         // leaves are sorted by x (triangle.cpp:5587-5591)
        for (int i = 0; i < n; i++) out[i] = xs[i];
        return;
    }
    const int divider = n >> 1;
    int* primary = axis ? ys : xs;
    int* other = axis ? xs : ys;
    for (int i = 0; i < divider; i++) side_[primary[i]] = 0;
    for (int i = divider; i < n; i++) side_[primary[i]] = 1;
    int* t = tmp_.data() + (other - (axis ? sx_.data() : sy_.data()));    // scratch window matching this range
    int l = 0, r = divider;
    for (int i = 0; i < n; i++) { const int id = other[i]; if (side_[id]) t[r++] = id; else t[l++] = id; }
    for (int i = 0; i < n; i++) other[i] = t[i];
    arrange(xs, ys, divider, 1 - axis, out);
    arrange(xs + divider, ys + divider, n - divider, 1 - axis, out + divider);
}

void Triangulator::merge(OTri& farleft, OTri& innerleft, OTri& innerright, OTri& farright, int axis)
{
    // mergehulls, triangle.cpp:5638-5934
    int ild = dest(innerleft), ila = apex(innerleft);
    int iro = org(innerright), ira = apex(innerright);

    if (axis == 1) {
        // horizontal cut: re-aim the four hull handles at the bottom-/top-most vertices (:5666-5704)
        int flp = org(farleft), fla = apex(farleft);
        int frp = dest(farright);
        while (y_[fla] < y_[flp]) {
            farleft = sym(lnext(farleft));
            flp = fla;
            fla = apex(farleft);
        }
        OTri check = sym(innerleft);
        int cv = apex(check);
        while (y_[cv] > y_[ild]) {
            innerleft = lnext(check);
            ila = ild;
            ild = cv;
            check = sym(innerleft);
            cv = apex(check);
        }
        while (y_[ira] < y_[iro]) {
            innerright = sym(lnext(innerright));
            iro = ira;
            ira = apex(innerright);
        }
        check = sym(farright);
        cv = apex(check);
        while (y_[cv] > y_[frp]) {
            farright = lnext(check);
            frp = cv;
            check = sym(farright);
            cv = apex(check);
        }
    }

    // lower common tangent (:5706-5726)
    for (bool changed = true; changed;) {
        changed = false;
        if (ccw(ild, ila, iro) > 0) {
            innerleft = sym(lprev(innerleft));
            ild = ila;
            ila = apex(innerleft);
            changed = true;
        }
        if (ccw(ira, iro, ild) > 0) {
            innerright = sym(lnext(innerright));
            iro = ira;
            ira = apex(innerright);
            changed = true;
        }
    }

    OTri leftcand = sym(innerleft), rightcand = sym(innerright);
    OTri base = make();                                  // bottom bounding triangle (:5731-5738)
    bond(base, innerleft);
    base = lnext(base);
    bond(base, innerright);
    base = lnext(base);
    set_org(base, iro);
    set_dest(base, ild);
    if (ild == org(farleft)) farleft = lnext(base);      // :5745-5752
    if (iro == dest(farright)) farright = lprev(base);

    int lowerleft = ild, lowerright = iro;
    int upperleft = apex(leftcand), upperright = apex(rightcand);

    for (;;) {
        const bool leftdone = ccw(upperleft, lowerleft, lowerright) <= 0;     // :5765-5768
        const bool rightdone = ccw(upperright, lowerleft, lowerright) <= 0;
        if (leftdone && rightdone) {
            OTri top = make();                           // top bounding triangle (:5771-5780)
            set_org(top, lowerleft);
            set_dest(top, lowerright);
            bond(top, base);
            top = lnext(top);
            bond(top, rightcand);
            top = lnext(top);
            bond(top, leftcand);
            if (axis == 1) {
                // restore the handles to the left-/right-most vertices (:5786-5809)
                int flp = org(farleft);
                int frp = dest(farright), fra = apex(farright);
                OTri check = sym(farleft);
                int cv = apex(check);
                while (x_[cv] < x_[flp]) {
                    farleft = lprev(check);
                    flp = cv;
                    check = sym(farleft);
                    cv = apex(check);
                }
                while (x_[fra] > x_[frp]) {
                    farright = sym(lprev(farright));
                    frp = fra;
                    fra = apex(farright);
                }
            }
            return;
        }
        if (!leftdone) {
            // flip away left-hull edges that are not Delaunay w.r.t. the knitting edge (:5813-5859)
            OTri next = sym(lprev(leftcand));
            int nextapex = apex(next);
            if (nextapex >= 0) {
                bool bad = incircle(lowerleft, lowerright, upperleft, nextapex) > 0;
                while (bad) {
                    next = lnext(next);
                    const OTri topcasing = sym(next);
                    next = lnext(next);
                    const OTri sidecasing = sym(next);
                    bond(next, topcasing);
                    bond(leftcand, sidecasing);
                    leftcand = lnext(leftcand);
                    const OTri outercasing = sym(leftcand);
                    next = lprev(next);
                    bond(next, outercasing);
                    set_org(leftcand, lowerleft);
                    set_dest(leftcand, -1);
                    set_apex(leftcand, nextapex);
                    set_org(next, -1);
                    set_dest(next, upperleft);
                    set_apex(next, nextapex);
                    upperleft = nextapex;
                    next = sidecasing;
                    nextapex = apex(next);
                    bad = nextapex >= 0 && incircle(lowerleft, lowerright, upperleft, nextapex) > 0;
                }
            }
        }
        if (!rightdone) {
            // same on the right hull (:5861-5907)
            OTri next = sym(lnext(rightcand));
            int nextapex = apex(next);
            if (nextapex >= 0) {
                bool bad = incircle(lowerleft, lowerright, upperright, nextapex) > 0;
                while (bad) {
                    next = lprev(next);
                    const OTri topcasing = sym(next);
                    next = lprev(next);
                    const OTri sidecasing = sym(next);
                    bond(next, topcasing);
                    bond(rightcand, sidecasing);
                    rightcand = lprev(rightcand);
                    const OTri outercasing = sym(rightcand);
                    next = lnext(next);
                    bond(next, outercasing);
                    set_org(rightcand, -1);
                    set_dest(rightcand, lowerright);
                    set_apex(rightcand, nextapex);
                    set_org(next, upperright);
                    set_dest(next, -1);
                    set_apex(next, nextapex);
                    upperright = nextapex;
                    next = sidecasing;
                    nextapex = apex(next);
                    bad = nextapex >= 0 && incircle(lowerleft, lowerright, upperright, nextapex) > 0;
                }
            }
        }
        // choose the next tooth; on co-circular quads the LEFT candidate wins (:5908-5910)
        if (leftdone || (!rightdone && incircle(upperleft, lowerleft, lowerright, upperright) > 0)) {
            bond(base, rightcand);
            base = lprev(rightcand);
            set_dest(base, lowerleft);
            lowerright = upperright;
            rightcand = sym(base);
            upperright = apex(rightcand);
        } else {
            bond(base, leftcand);
            base = lnext(leftcand);
            set_org(base, lowerright);
            lowerleft = upperleft;
            leftcand = sym(base);
            upperleft = apex(leftcand);
        }
    }
}

void Triangulator::recurse(int* s, int n, int axis, OTri& farleft, OTri& farright)
{
    // divconqrecurse, triangle.cpp:5953-6103
    if (n == 2) {
        farleft = make();
        set_org(farleft, s[0]); set_dest(farleft, s[1]);
        farright = make();
        set_org(farright, s[1]); set_dest(farright, s[0]);
        bond(farleft, farright);
        farleft = lprev(farleft); farright = lnext(farright);
        bond(farleft, farright);
        farleft = lprev(farleft); farright = lnext(farright);
        bond(farleft, farright);
        farleft = lprev(farright);
        return;
    }
    if (n == 3) {
        OTri mid = make(), t1 = make(), t2 = make(), t3 = make();
        const int area = ccw(s[0], s[1], s[2]);
        if (area == 0) {
            set_org(mid, s[0]); set_dest(mid, s[1]);
            set_org(t1, s[1]);  set_dest(t1, s[0]);
            set_org(t2, s[2]);  set_dest(t2, s[1]);
            set_org(t3, s[1]);  set_dest(t3, s[2]);
            bond(mid, t1); bond(t2, t3);
            mid = lnext(mid); t1 = lprev(t1); t2 = lnext(t2); t3 = lprev(t3);
            bond(mid, t3); bond(t1, t2);
            mid = lnext(mid); t1 = lprev(t1); t2 = lnext(t2); t3 = lprev(t3);
            bond(mid, t1); bond(t2, t3);
            farleft = t1;
            farright = t2;
        } else {
            const int p = area > 0 ? s[1] : s[2], q = area > 0 ? s[2] : s[1];
            set_org(mid, s[0]); set_dest(t1, s[0]); set_org(t3, s[0]);
            set_dest(mid, p);   set_org(t1, p);     set_dest(t2, p);
            set_apex(mid, q);   set_org(t2, q);     set_dest(t3, q);
            bond(mid, t1);
            mid = lnext(mid);
            bond(mid, t2);
            mid = lnext(mid);
            bond(mid, t3);
            t1 = lprev(t1); t2 = lnext(t2);
            bond(t1, t2);
            t1 = lprev(t1); t3 = lprev(t3);
            bond(t1, t3);
            t2 = lnext(t2); t3 = lprev(t3);
            bond(t2, t3);
            farleft = t1;
            farright = area > 0 ? t2 : lnext(farleft);
        }
        return;
    }
    const int divider = n >> 1;
    OTri innerleft, innerright;
    recurse(s, divider, 1 - axis, farleft, innerleft);
    recurse(s + divider, n - divider, 1 - axis, innerright, farright);
    merge(farleft, innerleft, innerright, farright, axis);
}

void Triangulator::run(const int32_t* x, const int32_t* y, int n, std::vector<int32_t>& out)
{
    out.clear();
    if (n < 2) return;
    x_ = x; y_ = y; seed_ = 1; ntri_ = 0;                 // randomseed reset per call, triangle.cpp:4030
    {
        int32_t xmin = x[0], xmax = x[0], ymin = y[0], ymax = y[0];
        for (int i = 1; i < n; i++) {
            xmin = std::min(xmin, x[i]); xmax = std::max(xmax, x[i]);
            ymin = std::min(ymin, y[i]); ymax = std::max(ymax, y[i]);
        }
        small_ = (int64_t)xmax - xmin < (1 << 14) && (int64_t)ymax - ymin < (1 << 14);
    }
    const size_t cap = 3 * (size_t)(3 * n + 8);
    if (nbr_.size() < cap) { nbr_.resize(cap); vtx_.resize(cap); }
    order_.resize(n);
    int* s = order_.data();
    int m = n;
    if (!presorted_order(n)) {
        for (int i = 0; i < n; i++) s[i] = i;
        sort(s, n);                                       // :6178
        m = 0;                                            // duplicates: keep the first (:6180-6196)
        for (int j = 1; j < n; j++)
            if (!(x[s[m]] == x[s[j]] && y[s[m]] == y[s[j]])) s[++m] = s[j];
        m++;
        const int divider = m >> 1;                       // :6197-6206
        if (m - divider >= 2) {
            if (divider >= 2) alternate(s, divider, 1);
            alternate(s + divider, m - divider, 1);
        }
    }
    if (m < 2) return;
    OTri hullleft, hullright;
    recurse(s, m, 0, hullleft, hullright);                // :6213
    // removeghosts (:6105-6148) frees exactly the triangles holding the ghost vertex; writeelements
    // (:7834-7853) then walks the pool in allocation order, corners (org, dest, apex) at orientation 0
    out.reserve((size_t)ntri_ * 3);
    for (int t = 0; t < ntri_; t++) {
        const int a = vtx_[3 * t + 1], b = vtx_[3 * t + 2], c = vtx_[3 * t];
        if (a < 0 || b < 0 || c < 0) continue;
        out.push_back(a); out.push_back(b); out.push_back(c);
    }
}

// =============================================================================================
// lattice filters, planes, raster records
// =============================================================================================
namespace {

// The three lattice filters run on a padded copy of the lattice (pitch Wc + 2*kPadC, kPadR rows above and
// below, padding = -1 = invalid) so that no neighbour access needs a bounds check and an 11-wide
// window row is two unaligned 128-bit loads.
constexpr int kPadC = 8, kPadR = 8;

// removeInconsistentSupportPoints, elas.cpp:174-209 (in place, u outer / v inner: every decision
// sees the invalidations made before it).  The reference counts all valid neighbours within the
// (2*win+1)^2 lattice window whose disparity differs by <= incon_threshold and invalidates the
// point when the count is below incon_min_support; only that comparison is observable, so the
// count stops as soon as it reaches incon_min_support (rows nearest the centre are visited first).
// A window row (<= 16 lattice points) is compared with SSE2: lanes d2 in [max(d-thr,0), d+thr].
// `cells` lists the offsets (into the padded array) of the cells that are valid on entry, in the
// reference's scan order (u outer, v inner); a filter only ever invalidates the cell it is visiting, so
// the cells to visit are exactly those and the 65 % invalid cells cost nothing.
void remove_inconsistent_padded(const elas_b200_params& p, int16_t* P, int pitch, const std::vector<int32_t>& cells)
{
    const int win = p.incon_window_size, need = p.incon_min_support, thr = p.incon_threshold;
    const int span = 2 * win + 1;                           // lanes of a window row, <= 16 on this path
    const uint32_t lane_mask = span >= 16 ? 0xFFFFFFFFu : ((1u << (2 * span)) - 1u);
    for (const int32_t off : cells) {
        int16_t* centre = P + off;
        const int d = *centre;
        const __m128i lo1 = _mm_set1_epi16((short)(std::max(d - thr, 0) - 1));      // x > lo-1
        const __m128i hi1 = _mm_set1_epi16((short)std::min(d + thr + 1, 32767));    // x < hi+1
        int support = 0;
        for (int k = 0; k <= 2 * win && support < need; k++) {
            const int dv = (k & 1) ? (k + 1) / 2 : -(k / 2);              // v, v+1, v-1, v+2, v-2, ...
            const int16_t* row = centre + (ptrdiff_t)dv * pitch - win;
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(row));
            const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(row + 8));
            const __m128i ma = _mm_and_si128(_mm_cmpgt_epi16(a, lo1), _mm_cmpgt_epi16(hi1, a));
            const __m128i mb = _mm_and_si128(_mm_cmpgt_epi16(b, lo1), _mm_cmpgt_epi16(hi1, b));
            const uint32_t m = ((uint32_t)_mm_movemask_epi8(ma) | ((uint32_t)_mm_movemask_epi8(mb) << 16)) & lane_mask;
            support += __builtin_popcount(m) >> 1;
        }
        if (support < need) *centre = -1;
    }
}

// generic window sizes (incon_window_size > 7): the same scan, scalar
void remove_inconsistent_scalar(const elas_b200_params& p, int16_t* D, int Wc, int Hc)
{
    const int win = p.incon_window_size, need = p.incon_min_support, thr = p.incon_threshold;
    for (int u = 0; u < Wc; u++) {
        const int u_lo = std::max(u - win, 0), u_hi = std::min(u + win, Wc - 1);
        for (int v = 0; v < Hc; v++) {
            const int d = D[v * Wc + u];
            if (d < 0) continue;
            const int lo = std::max(d - thr, 0), hi = d + thr;
            int support = 0;
            for (int k = 0; k <= 2 * win && support < need; k++) {
                const int v2 = v + ((k & 1) ? (k + 1) / 2 : -(k / 2));
                if (v2 < 0 || v2 >= Hc) continue;
                const int16_t* row = D + v2 * Wc;
                for (int u2 = u_lo; u2 <= u_hi; u2++) support += row[u2] >= lo && row[u2] <= hi;
            }
            if (support < need) D[v * Wc + u] = -1;
        }
    }
}

// removeRedundantSupportPoints, elas.cpp:213-279 (in place): a point goes when, in BOTH directions along
// the axis, a valid point with |d - d2| <= thresh lies within max_dist lattice steps.  `step` is the
// element stride of the axis in the padded array (pitch for the vertical pass, 1 for the horizontal).
void remove_redundant_padded(int16_t* P, const std::vector<int32_t>& cells, int max_dist, int thresh, ptrdiff_t step)
{
    for (const int32_t off : cells) {
        int16_t* centre = P + off;
        const int d = *centre;
        if (d < 0) continue;                                  // invalidated by an earlier filter
        const int lo = std::max(d - thresh, 0), hi = d + thresh;
        bool back = false, fwd = false;
        for (int j = 1; j <= max_dist; j++) { const int x = centre[-j * step]; back |= x >= lo && x <= hi; }
        if (!back) continue;
        for (int j = 1; j <= max_dist; j++) { const int x = centre[j * step]; fwd |= x >= lo && x <= hi; }
        if (fwd) *centre = -1;
    }
}

// removeRedundantSupportPoints without padding (max_dist beyond the padding)
void remove_redundant_scalar(int16_t* D, int Wc, int Hc, int max_dist, int thresh, bool vertical)
{
    const int du = vertical ? 0 : 1, dv = vertical ? 1 : 0;
    for (int u = 0; u < Wc; u++)
        for (int v = 0; v < Hc; v++) {
            const int d = D[v * Wc + u];
            if (d < 0) continue;
            bool redundant = true;
            for (int dir = -1; dir <= 1 && redundant; dir += 2) {
                bool support = false;
                int u2 = u, v2 = v;
                for (int j = 0; j < max_dist; j++) {
                    u2 += dir * du; v2 += dir * dv;
                    if (u2 < 0 || v2 < 0 || u2 >= Wc || v2 >= Hc) break;
                    const int d2 = D[v2 * Wc + u2];
                    if (d2 >= 0 && std::abs(d - d2) <= thresh) { support = true; break; }
                }
                if (!support) redundant = false;
            }
            if (redundant) D[v * Wc + u] = -1;
        }
}

// addCornerSupportPoints, elas.cpp:283-318
void add_corners(int W, int H, std::vector<int32_t>& sup)
{
    const int n = (int)sup.size() / 3;
    int32_t b[6][3] = {{0, 0, 0}, {0, H - 1, 0}, {W - 1, 0, 0}, {W - 1, H - 1, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int i = 0; i < 4; i++) {
        int best = 10000000;
        for (int j = 0; j < n; j++) {
            const int du = b[i][0] - sup[3 * j], dv = b[i][1] - sup[3 * j + 1];
            const int dist = du * du + dv * dv;
            if (dist < best) { best = dist; b[i][2] = sup[3 * j + 2]; }
        }
    }
    for (int k = 0; k < 2; k++) { b[4 + k][0] = b[2 + k][0] + b[2 + k][2]; b[4 + k][1] = b[2 + k][1]; b[4 + k][2] = b[2 + k][2]; }
    for (auto& r : b) sup.insert(sup.end(), r, r + 3);
}

// Matrix::solve for a 3x3 system with one right-hand side (matrix.cpp:414-502): Gauss-Jordan, full
// pivoting, double precision; false when a pivot is below 1e-20.
bool solve3(double A[3][3], double b[3])
{
    int ipiv[3] = {0, 0, 0};
    for (int i = 0; i < 3; i++) {
        double big = 0.0;
        int irow = 0, icol = 0;
        for (int j = 0; j < 3; j++)
            if (ipiv[j] != 1)
                for (int k = 0; k < 3; k++)
                    if (ipiv[k] == 0 && std::fabs(A[j][k]) >= big) { big = std::fabs(A[j][k]); irow = j; icol = k; }
        ++ipiv[icol];
        if (irow != icol) {
            for (int l = 0; l < 3; l++) std::swap(A[irow][l], A[icol][l]);
            std::swap(b[irow], b[icol]);
        }
        if (std::fabs(A[icol][icol]) < 1e-20) return false;
        const double pivinv = 1.0 / A[icol][icol];
        A[icol][icol] = 1.0;
        for (int l = 0; l < 3; l++) A[icol][l] *= pivinv;
        b[icol] *= pivinv;
        for (int ll = 0; ll < 3; ll++)
            if (ll != icol) {
                const double dum = A[ll][icol];
                A[ll][icol] = 0.0;
                for (int l = 0; l < 3; l++) A[ll][l] -= A[icol][l] * dum;
                b[ll] -= b[icol] * dum;
            }
    }
    return true;
}

// computeDisparityPlanes, elas.cpp:605-680
void disparity_planes(const std::vector<int32_t>& sup, const std::vector<int32_t>& tri, std::vector<float>& planes)
{
    const size_t nt = tri.size() / 3;
    planes.resize(nt * 6);
    for (size_t i = 0; i < nt; i++)
        for (int k = 0; k < 2; k++) {
            double A[3][3], b[3];
            for (int c = 0; c < 3; c++) {
                const int32_t* s = &sup[3 * (size_t)tri[3 * i + c]];
                A[c][0] = k ? s[0] - s[2] : s[0];
                A[c][1] = s[1];
                A[c][2] = 1;
                b[c] = s[2];
            }
            float* o = &planes[6 * i + 3 * k];
            if (solve3(A, b)) { o[0] = (float)b[0]; o[1] = (float)b[1]; o[2] = (float)b[2]; }
            else o[0] = o[1] = o[2] = 0.f;
        }
}

// per-triangle set-up of computeDisparity, elas.cpp:1006-1072
void raster_records(const std::vector<int32_t>& sup, const std::vector<int32_t>& tri,
                    const std::vector<float>& planes, int right_image, std::vector<TriRaster>& out)
{
    const size_t nt = tri.size() / 3;
    out.resize(nt);
    for (size_t i = 0; i < nt; i++) {
        const float* pl = &planes[6 * i];
        TriRaster r{};
        const float pd = right_image ? pl[0] : pl[3];
        r.pa = right_image ? pl[3] : pl[0];
        r.pb = right_image ? pl[4] : pl[1];
        r.pc = right_image ? pl[5] : pl[2];
        float tu[3], tv[3];
        for (int k = 0; k < 3; k++) {
            const int32_t* s = &sup[3 * (size_t)tri[3 * i + k]];
            tu[k] = right_image ? (float)(s[0] - s[2]) : (float)s[0];
            tv[k] = (float)s[1];
        }
        for (int j = 0; j < 3; j++)                       // :1043-1053
            for (int k = 0; k < j; k++)
                if (tu[k] > tu[j]) { std::swap(tu[j], tu[k]); std::swap(tv[j], tv[k]); }
        const float Au = tu[0], Av = tv[0], Bu = tu[1], Bv = tv[1], Cu = tu[2], Cv = tv[2];
        float ABa = 0, ACa = 0, BCa = 0;                  // :1061-1067
        if ((int32_t)Au != (int32_t)Bu) ABa = (Av - Bv) / (Au - Bu);
        if ((int32_t)Au != (int32_t)Cu) ACa = (Av - Cv) / (Au - Cu);
        if ((int32_t)Bu != (int32_t)Cu) BCa = (Bv - Cv) / (Bu - Cu);
        r.ABa = ABa; r.ACa = ACa; r.BCa = BCa;
        r.ABb = Av - ABa * Au;
        r.ACb = Av - ACa * Au;
        r.BCb = Bv - BCa * Bu;
        r.uA = (int32_t)Au; r.uB = (int32_t)Bu; r.uC = (int32_t)Cu;
        r.valid = (std::fabs(r.pa) < 0.7 && std::fabs(pd) < 0.7) ? 2 : 0;   // :1072 (float |.| compared in double); bit 1 of K7's packed pixel state
        out[i] = r;
    }
}

}  // namespace

int HostStage::run(const FrameGeom& g, const elas_b200_params& p, int16_t* dcan, bool keep_stages, bool with_planes)
{
    const int Wc = g.Wc, Hc = g.Hc;
    support.clear();
    if (p.incon_window_size <= 7) {
        // padded working copy (see kPadC/kPadR)
        const int pitch = Wc + 2 * kPadC;
        pad_.assign((size_t)pitch * (Hc + 2 * kPadR) + 16, (int16_t)-1);
        int16_t* P = pad_.data();
        for (int v = 0; v < Hc; v++) std::memcpy(P + (size_t)(v + kPadR) * pitch + kPadC, dcan + (size_t)v * Wc, (size_t)Wc * 2);
        auto unpad = [&](int16_t* dst) {
            for (int v = 0; v < Hc; v++) std::memcpy(dst + (size_t)v * Wc, P + (size_t)(v + kPadR) * pitch + kPadC, (size_t)Wc * 2);
        };
        // valid cells in scan order (u outer, v inner): rows are read contiguously, counted per column,
        // and scattered into the prefix-summed column-major layout
        col_fill_.assign((size_t)Wc + 1, 0);
        for (int v = 0; v < Hc; v++) {
            const int16_t* row = dcan + (size_t)v * Wc;
            for (int u = 0; u < Wc; u++) col_fill_[u + 1] += row[u] >= 0;
        }
        for (int u = 1; u <= Wc; u++) col_fill_[u] += col_fill_[u - 1];
        cells_.resize((size_t)col_fill_[Wc]);
        for (int v = 0; v < Hc; v++) {
            const int16_t* row = dcan + (size_t)v * Wc;
            const int32_t base = (v + kPadR) * pitch + kPadC;
            for (int u = 0; u < Wc; u++)
                if (row[u] >= 0) cells_[col_fill_[u]++] = base + u;
        }
        remove_inconsistent_padded(p, P, pitch, cells_);              // elas.cpp:496
        if (keep_stages) { dcan_incon.resize((size_t)Wc * Hc); unpad(dcan_incon.data()); }
        remove_redundant_padded(P, cells_, 5, 1, pitch);              // :501 (vertical)
        remove_redundant_padded(P, cells_, 5, 1, 1);                  // :502 (horizontal)
        unpad(dcan);
        // :505-517: the survivors outside lattice row 0 / column 0, already in u-outer / v-inner order
        for (const int32_t off : cells_) {
            const int d = P[off];
            if (d < 0) continue;
            const int vc = off / pitch - kPadR, uc = off - (vc + kPadR) * pitch - kPadC;
            if (uc < 1 || vc < 1) continue;
            support.push_back(uc * g.step); support.push_back(vc * g.step); support.push_back(d);
        }
    } else {
        remove_inconsistent_scalar(p, dcan, Wc, Hc);
        if (keep_stages) dcan_incon.assign(dcan, dcan + (size_t)Wc * Hc);
        remove_redundant_scalar(dcan, Wc, Hc, 5, 1, true);
        remove_redundant_scalar(dcan, Wc, Hc, 5, 1, false);
        for (int uc = 1; uc < Wc; uc++)                               // :505-517, u outer / v inner
            for (int vc = 1; vc < Hc; vc++) {
                const int d = dcan[vc * Wc + uc];
                if (d >= 0) { support.push_back(uc * g.step); support.push_back(vc * g.step); support.push_back(d); }
            }
    }
    if (p.add_corners) add_corners(g.W, g.H, support);            // :520-523
    n_support = (int)support.size() / 3;
    for (int k = 0; k < 2; k++) { tri[k].clear(); planes[k].clear(); raster[k].clear(); }
    // units is grown geometrically and trimmed to its final size once (no per-triangle zero fill)
    size_t n_units_ints = 0;
    if (n_support < 3) { units.clear(); return n_support; }       // :69-75

    px_.resize(n_support); py_.resize(n_support);
    for (int k = 0; k < 2; k++) {                                 // :80-81, :534-559
        for (int i = 0; i < n_support; i++) {
            px_[i] = k ? support[3 * i] - support[3 * i + 2] : support[3 * i];
            py_[i] = support[3 * i + 1];
        }
        delaunay_.run(px_.data(), py_.data(), n_support, tri[k]);
        // work units: 32-column chunks x kRasterBandRows-row bands of each triangle's bounding box
        const size_t nt = tri[k].size() / 3;
        const int32_t* tk = tri[k].data();
        for (size_t t = 0; t < nt; t++) {
            const int a = tk[3 * t], b = tk[3 * t + 1], c = tk[3 * t + 2];
            const int u_lo = std::max(std::min(px_[a], std::min(px_[b], px_[c])), 0);
            const int u_hi = std::min(std::max(px_[a], std::max(px_[b], px_[c])), g.W);      // columns [u_lo, u_hi)
            // one row of slack below the smallest corner row: an edge line evaluated in float may truncate to it
            const int v_lo = std::max(std::min(py_[a], std::min(py_[b], py_[c])) - 1, 0);
            const int v_hi = std::min(std::max(py_[a], std::max(py_[b], py_[c])) + 1, g.H);  // rows [v_lo, v_hi)
            const int chunks = (u_hi - u_lo + 31) / 32, bands = (v_hi - v_lo + kRasterBandRows - 1) / kRasterBandRows;
            if (chunks <= 0 || bands <= 0) continue;
            const size_t need = 2 * (size_t)chunks * bands;
            if (n_units_ints + need > units.size()) units.resize(std::max(2 * units.size(), n_units_ints + need + 4096));
            int32_t* o = units.data() + n_units_ints;
            for (int ch = 0; ch < chunks; ch++)
                for (int bd = 0; bd < bands; bd++) {
                    *o++ = (int32_t)t | (k << 30);
                    *o++ = ch | (bd << 16);
                }
            n_units_ints += need;
        }
        if (with_planes) {
            disparity_planes(support, tri[k], planes[k]);         // :87-88
            raster_records(support, tri[k], planes[k], k, raster[k]);
        }
    }
    units.resize(n_units_ints);
    return n_support;
}

}  // namespace elasb
