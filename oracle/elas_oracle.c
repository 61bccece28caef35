/*
 * TEST INFRASTRUCTURE -- not part of the product.
 *
 * elas_oracle.c: scalar, plain-C restatement of the dense-stereo hot path that the reference's
 * Elas::process runs per frame (libelas/src/elas.cpp:32-170 and what it calls in descriptor.cpp,
 * filter.cpp, matrix.cpp and triangle.cpp).  It exists only to check the CUDA path: tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg are its only callers.  The product
 * (stereo-vision_b200/) never links, imports or executes it.
 *
 * Parity pinning: the reference holds no tests or golden vectors for this path, so
 * this restatement is pinned against the reference itself: oracle/_ref/libelas_ref.so (the
 * unmodified sources compiled by oracle/Makefile) must agree with it bit for bit, stage by stage
 * (tests/test_oracle.py: seeded synthetic pairs, the seven `./elas demo` pairs of libelas/img,
 * the 4096x2160 configuration and 120 degenerate point sets for the triangulator), and against vectors generated from that build and
 * committed under tests/golden/ (tests/golden/make_golden.py).
 *
 * Every function cites the reference lines it restates.  No SIMD, no threads, one frame at a
 * time.  Floating point follows the reference's x86-64 SSE evaluation (no contraction: build with
 * -ffp-contract=off), so float results are bit-identical, not merely close.
 *
 * Defined behaviour where the reference's is undefined: descriptor
 * pixels outside v in [3,H-3), u in [3,W-3) are 0.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "elas_b200.h"

#define MINI(a, b) ((a) < (b) ? (a) : (b))
#define MAXI(a, b) ((a) > (b) ? (a) : (b))

/* ------------------------------------------------------------------------------------------ */
/* stage store                                                                                */
/* ------------------------------------------------------------------------------------------ */

#define MAX_STAGES 64
static struct { char name[24]; void* data; int64_t bytes; } g_stage[MAX_STAGES];
static int g_nstage = 0;
static int g_keep = 0;

static void stage_clear(void)
{
    for (int i = 0; i < g_nstage; i++) free(g_stage[i].data);
    g_nstage = 0;
}

static void keep(const char* name, const void* data, int64_t bytes)
{
    if (!g_keep || g_nstage >= MAX_STAGES) return;
    strncpy(g_stage[g_nstage].name, name, sizeof g_stage[g_nstage].name - 1);
    g_stage[g_nstage].name[sizeof g_stage[g_nstage].name - 1] = 0;
    g_stage[g_nstage].data = malloc(bytes > 0 ? (size_t)bytes : 1);
    memcpy(g_stage[g_nstage].data, data, (size_t)bytes);
    g_stage[g_nstage].bytes = bytes;
    g_nstage++;
}

int64_t oracle_stage_bytes(const char* name)
{
    for (int i = 0; i < g_nstage; i++)
        if (!strcmp(g_stage[i].name, name)) return g_stage[i].bytes;
    return -1;
}

int32_t oracle_stage_read(const char* name, void* dst, int64_t cap)
{
    for (int i = 0; i < g_nstage; i++)
        if (!strcmp(g_stage[i].name, name)) {
            if (g_stage[i].bytes > cap) return ELAS_B200_E_BAD_ARG;
            memcpy(dst, g_stage[i].data, (size_t)g_stage[i].bytes);
            return 0;
        }
    return ELAS_B200_E_NO_STAGE;
}

/* ------------------------------------------------------------------------------------------ */
/* a3 + a4: Sobel responses and the 16-byte descriptor                                        */
/* ------------------------------------------------------------------------------------------ */

static uint8_t sat_u8(int x) { return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x)); }

/*
 * filter.cpp:408-416 (sobel3x3) = convolve_cols_3x3 (:372-405) then the two row passes
 * (:227-267, :176-222), all in int16 with an arithmetic >>2, +128 and unsigned saturation:
 *   S(u,v) = I(u,v-1) + 2 I(u,v) + I(u,v+1)          T(u,v) = I(u,v-1) - I(u,v+1)
 *   du(u,v) = sat8(((S(u-1,v) - S(u+1,v)) >> 2) + 128)
 *   dv(u,v) = sat8(((T(u-1,v) + 2 T(u,v) + T(u+1,v)) >> 2) + 128)
 * The reference indexes the padded image flat; for rows 1..H-2 and columns 1..bpl-2 (all the
 * descriptor ever gathers) flat and 2-D indexing coincide.
 * descriptor.cpp:48-121 (createDescriptor) gathers 12 du and 4 dv taps per pixel.
 */
static void descriptor(const uint8_t* I, int W, int H, int bpl, int half, uint8_t* desc)
{
    uint8_t* du = (uint8_t*)calloc((size_t)bpl * H, 1);
    uint8_t* dv = (uint8_t*)calloc((size_t)bpl * H, 1);
    for (int v = 1; v < H - 1; v++)
        for (int u = 1; u < bpl - 1; u++) {
            const uint8_t* r0 = I + (size_t)(v - 1) * bpl;
            const uint8_t* r1 = I + (size_t)v * bpl;
            const uint8_t* r2 = I + (size_t)(v + 1) * bpl;
            int Sl = r0[u - 1] + 2 * r1[u - 1] + r2[u - 1];
            int Sr = r0[u + 1] + 2 * r1[u + 1] + r2[u + 1];
            int Tl = r0[u - 1] - r2[u - 1], Tc = r0[u] - r2[u], Tr = r0[u + 1] - r2[u + 1];
            du[(size_t)v * bpl + u] = sat_u8(((Sl - Sr) >> 2) + 128);
            dv[(size_t)v * bpl + u] = sat_u8(((Tl + 2 * Tc + Tr) >> 2) + 128);
        }
    memset(desc, 0, (size_t)16 * W * H);
    /* descriptor.cpp:54 (half resolution: v=4,6,..) and :88 (full: v=3..H-4) */
    for (int v = half ? 4 : 3; v < H - 3; v += half ? 2 : 1) {
        const uint8_t *u0 = du + (size_t)(v - 2) * bpl, *u1 = du + (size_t)(v - 1) * bpl,
                      *u2 = du + (size_t)v * bpl, *u3 = du + (size_t)(v + 1) * bpl,
                      *u4 = du + (size_t)(v + 2) * bpl;
        const uint8_t *v1 = dv + (size_t)(v - 1) * bpl, *v2 = dv + (size_t)v * bpl,
                      *v3 = dv + (size_t)(v + 1) * bpl;
        for (int u = 3; u < W - 3; u++) {
            uint8_t* o = desc + ((size_t)v * W + u) * 16;
            o[0] = u0[u];      o[1] = u1[u - 2];  o[2] = u1[u];      o[3] = u1[u + 2];
            o[4] = u2[u - 1];  o[5] = u2[u];      o[6] = u2[u];      o[7] = u2[u + 1];
            o[8] = u3[u - 2];  o[9] = u3[u];      o[10] = u3[u + 2]; o[11] = u4[u];
            o[12] = v1[u];     o[13] = v2[u - 1]; o[14] = v2[u + 1]; o[15] = v3[u];
        }
    }
    free(du); free(dv);
}

static int sad16(const uint8_t* a, const uint8_t* b)
{
    int s = 0;
    for (int i = 0; i < 16; i++) s += abs((int)a[i] - (int)b[i]);
    return s;
}

static int texture16(const uint8_t* a)
{
    int s = 0;
    for (int i = 0; i < 16; i++) s += abs((int)a[i] - 128);
    return s;
}

/* ------------------------------------------------------------------------------------------ */
/* a5: support match at one lattice point (elas.cpp:322-445)                                  */
/* ------------------------------------------------------------------------------------------ */

static int matching_disparity(const elas_b200_params* p, int W, int H, int u, int v,
                              const uint8_t* desc1, const uint8_t* desc2, int right_image)
{
    const int u_step = 2, v_step = 2, window = 3;                       /* :325-327 */
    if (!(u >= window + u_step && u <= W - window - 1 - u_step &&
          v >= window + v_step && v <= H - window - 1 - v_step)) return -1;   /* :337 */
    const uint8_t* own   = right_image ? desc2 : desc1;                  /* :342-351 */
    const uint8_t* other = right_image ? desc1 : desc2;
    if (texture16(own + ((size_t)v * W + u) * 16) < p->support_texture) return -1;  /* :358-366 */

    int dmin = MAXI(p->disp_min, 0);                                     /* :384-387 */
    int dmax = right_image ? MINI(p->disp_max, W - u - window - u_step)
                           : MINI(p->disp_max, u - window - u_step);
    if (dmax - dmin < 10) return -1;                                     /* :390 */

    int e1 = 32767, d1 = -1, e2 = 32767, d2 = -1;                        /* :378-381 */
    for (int d = dmin; d <= dmax; d++) {                                 /* :396-429 */
        int uw = right_image ? u + d : u - d;
        int sum = 0;
        for (int by = -1; by <= 1; by += 2)
            for (int bx = -1; bx <= 1; bx += 2) {                        /* the four blocks :329-332 */
                size_t a = ((size_t)(v + by * v_step) * W + (u + bx * u_step)) * 16;
                size_t b = ((size_t)(v + by * v_step) * W + (uw + bx * u_step)) * 16;
                sum += sad16(own + a, other + b);
            }
        if (sum < e1) { e2 = e1; d2 = d1; e1 = sum; d1 = d; }
        else if (sum < e2) { e2 = sum; d2 = d; }
    }
    if (d1 >= 0 && d2 >= 0 && (float)e1 < p->support_threshold * (float)e2) return d1;   /* :432 */
    return -1;
}

/* a7: elas.cpp:174-209, in place, u outer / v inner */
static void remove_inconsistent(const elas_b200_params* p, int16_t* D, int Wc, int Hc)
{
    int win = p->incon_window_size;
    for (int u = 0; u < Wc; u++)
        for (int v = 0; v < Hc; v++) {
            int d = D[v * Wc + u];
            if (d < 0) continue;
            int support = 0;
            for (int u2 = u - win; u2 <= u + win; u2++)
                for (int v2 = v - win; v2 <= v + win; v2++)
                    if (u2 >= 0 && v2 >= 0 && u2 < Wc && v2 < Hc) {
                        int d2 = D[v2 * Wc + u2];
                        if (d2 >= 0 && abs(d - d2) <= p->incon_threshold) support++;
                    }
            if (support < p->incon_min_support) D[v * Wc + u] = -1;
        }
}

/* a8: elas.cpp:213-279, in place, u outer / v inner */
static void remove_redundant(int16_t* D, int Wc, int Hc, int max_dist, int thresh, int vertical)
{
    int du[2] = {0, 0}, dv[2] = {0, 0};
    if (vertical) { dv[0] = -1; dv[1] = 1; } else { du[0] = -1; du[1] = 1; }
    for (int u = 0; u < Wc; u++)
        for (int v = 0; v < Hc; v++) {
            int d = D[v * Wc + u];
            if (d < 0) continue;
            int redundant = 1;
            for (int i = 0; i < 2 && redundant; i++) {
                int u2 = u, v2 = v, support = 0;
                for (int j = 0; j < max_dist; j++) {
                    u2 += du[i]; v2 += dv[i];
                    if (u2 < 0 || v2 < 0 || u2 >= Wc || v2 >= Hc) break;
                    int d2 = D[v2 * Wc + u2];
                    if (d2 >= 0 && abs(d - d2) <= thresh) { support = 1; break; }
                }
                if (!support) redundant = 0;
            }
            if (redundant) D[v * Wc + u] = -1;
        }
}

/* a9: elas.cpp:283-318 */
static int add_corner_points(int W, int H, int32_t* sup, int n)
{
    int32_t b[6][3] = {{0, 0, 0}, {0, H - 1, 0}, {W - 1, 0, 0}, {W - 1, H - 1, 0}};
    for (int i = 0; i < 4; i++) {
        int best = 10000000;
        for (int j = 0; j < n; j++) {
            int du = b[i][0] - sup[3 * j], dv = b[i][1] - sup[3 * j + 1];
            int dist = du * du + dv * dv;
            if (dist < best) { best = dist; b[i][2] = sup[3 * j + 2]; }
        }
    }
    b[4][0] = b[2][0] + b[2][2]; b[4][1] = b[2][1]; b[4][2] = b[2][2];
    b[5][0] = b[3][0] + b[3][2]; b[5][1] = b[3][1]; b[5][2] = b[3][2];
    for (int i = 0; i < 6; i++) memcpy(sup + 3 * (n + i), b[i], sizeof b[i]);
    return n + 6;
}

/*
 * a6: elas.cpp:449-530.  Lattice D_can is calloc'ed (:464) so row 0 and column 0 hold the VALID
 * disparity 0 while the filters run and are excluded only when points are collected.
 * `sup` must hold room for (Wc*Hc+6) triples.  Returns the number of support points.
 */
static int support_matches(const elas_b200_params* p, int W, int H,
                           const uint8_t* desc1, const uint8_t* desc2, int32_t* sup,
                           int* Wc_out, int* Hc_out)
{
    int step = p->candidate_stepsize;
    if (p->subsampling) step += step % 2;                                /* :453-457 */
    int Wc = 0, Hc = 0;
    for (int u = 0; u < W; u += step) Wc++;
    for (int v = 0; v < H; v += step) Hc++;
    int16_t* D = (int16_t*)calloc((size_t)Wc * Hc, sizeof(int16_t));
    for (int uc = 1; uc < Wc; uc++)
        for (int vc = 1; vc < Hc; vc++) {
            int u = uc * step, v = vc * step;
            D[vc * Wc + uc] = -1;
            int d = matching_disparity(p, W, H, u, v, desc1, desc2, 0);
            if (d >= 0) {
                int d2 = matching_disparity(p, W, H, u - d, v, desc1, desc2, 1);
                if (d2 >= 0 && abs(d - d2) <= p->lr_threshold) D[vc * Wc + uc] = (int16_t)d;
            }
        }
    keep("dcan_raw", D, (int64_t)Wc * Hc * 2);
    remove_inconsistent(p, D, Wc, Hc);                                    /* :496 */
    keep("dcan_incon", D, (int64_t)Wc * Hc * 2);
    remove_redundant(D, Wc, Hc, 5, 1, 1);                                 /* :501 */
    remove_redundant(D, Wc, Hc, 5, 1, 0);                                 /* :502 */
    keep("dcan", D, (int64_t)Wc * Hc * 2);
    int32_t lat[2] = {Wc, Hc};
    keep("lattice_dims", lat, sizeof lat);
    int n = 0;
    for (int uc = 1; uc < Wc; uc++)                                       /* :505-517 */
        for (int vc = 1; vc < Hc; vc++)
            if (D[vc * Wc + uc] >= 0) {
                sup[3 * n] = uc * step; sup[3 * n + 1] = vc * step; sup[3 * n + 2] = D[vc * Wc + uc];
                n++;
            }
    if (p->add_corners) n = add_corner_points(W, H, sup, n);              /* :520-523 */
    free(D);
    *Wc_out = Wc; *Hc_out = Hc;
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* a10: Delaunay triangulation, restating Triangle 1.6's divide-and-conquer with alternating   */
/* cuts as the reference runs it with switches "zQB" (elas.cpp:581; triangle.cpp:6160-6217).    */
/*                                                                                            */
/* Same triangle-based structure with ghost ("bounding") triangles, same allocation order, same */
/* strict tie-breaks, so that triangle corners AND output order equal writeelements()           */
/* (triangle.cpp:7800-7853).  The adaptive float predicates (triangle.cpp:2706, :3334) return    */
/* exact signs; coordinates are integers, so exact integer determinants give the same signs.    */
/* ------------------------------------------------------------------------------------------ */

typedef struct { int t, o; } otri;          /* oriented triangle: index + orientation 0..2 */

typedef struct {
    int* nbr;        /* [3*t+o] = (neighbour index << 2) | neighbour orientation, -1 = none */
    int* vtx;        /* [3*t+o] = vertex id, -1 = the ghost vertex (NULL in Triangle)        */
    int ntri;
    const int32_t *x, *y;
    uint64_t seed;   /* triangle.cpp:550, reset to 1 per call (:4030) */
} dmesh;

static const int plus1[3] = {1, 2, 0}, minus1[3] = {2, 0, 1};     /* triangle.cpp:812-813 */

static otri d_sym(const dmesh* m, otri a) { int e = m->nbr[3 * a.t + a.o]; otri r = {e >> 2, e & 3}; return r; }
static otri d_lnext(otri a) { otri r = {a.t, plus1[a.o]}; return r; }
static otri d_lprev(otri a) { otri r = {a.t, minus1[a.o]}; return r; }
static int d_org(const dmesh* m, otri a) { return m->vtx[3 * a.t + plus1[a.o]]; }
static int d_dest(const dmesh* m, otri a) { return m->vtx[3 * a.t + minus1[a.o]]; }
static int d_apex(const dmesh* m, otri a) { return m->vtx[3 * a.t + a.o]; }
static void d_setorg(dmesh* m, otri a, int v) { m->vtx[3 * a.t + plus1[a.o]] = v; }
static void d_setdest(dmesh* m, otri a, int v) { m->vtx[3 * a.t + minus1[a.o]] = v; }
static void d_setapex(dmesh* m, otri a, int v) { m->vtx[3 * a.t + a.o] = v; }
static void d_bond(dmesh* m, otri a, otri b)
{
    m->nbr[3 * a.t + a.o] = (b.t << 2) | b.o;
    m->nbr[3 * b.t + b.o] = (a.t << 2) | a.o;
}
static otri d_make(dmesh* m)                                       /* maketriangle, :2201-2229 */
{
    int t = m->ntri++;
    for (int i = 0; i < 3; i++) { m->nbr[3 * t + i] = -1; m->vtx[3 * t + i] = -1; }
    otri r = {t, 0};
    return r;
}

/* sign of counterclockwise(pa,pb,pc), triangle.cpp:2706-2745 */
static int d_ccw(const dmesh* m, int a, int b, int c)
{
    int64_t l = (int64_t)(m->x[a] - m->x[c]) * (m->y[b] - m->y[c]);
    int64_t r = (int64_t)(m->y[a] - m->y[c]) * (m->x[b] - m->x[c]);
    return (l > r) - (l < r);
}

/* sign of incircle(pa,pb,pc,pd), triangle.cpp:3334-3380 */
static int d_incircle(const dmesh* m, int a, int b, int c, int d)
{
    int64_t adx = m->x[a] - m->x[d], ady = m->y[a] - m->y[d];
    int64_t bdx = m->x[b] - m->x[d], bdy = m->y[b] - m->y[d];
    int64_t cdx = m->x[c] - m->x[d], cdy = m->y[c] - m->y[d];
    __int128 det = (__int128)(adx * adx + ady * ady) * (bdx * cdy - cdx * bdy)
                 + (__int128)(bdx * bdx + bdy * bdy) * (cdx * ady - adx * cdy)
                 + (__int128)(cdx * cdx + cdy * cdy) * (adx * bdy - bdx * ady);
    return (det > 0) - (det < 0);
}

static int d_random(dmesh* m, unsigned choices)                     /* randomnation, :4045-4049 */
{
    m->seed = (m->seed * 1366u + 150889u) % 714025u;
    return (int)(m->seed / (714025u / choices + 1));
}

/* lexicographic "a before b" with primary axis `axis` (0: x then y, 1: y then x) */
static int d_less(const dmesh* m, int a, int b, int axis)
{
    int32_t a1 = axis ? m->y[a] : m->x[a], a2 = axis ? m->x[a] : m->y[a];
    int32_t b1 = axis ? m->y[b] : m->x[b], b2 = axis ? m->x[b] : m->y[b];
    return a1 < b1 || (a1 == b1 && a2 < b2);
}

/* vertexsort, triangle.cpp:5446-5499: randomised quicksort on (x, y) */
static void d_sort(dmesh* m, int* s, int n)
{
    if (n == 2) {
        if (d_less(m, s[1], s[0], 0)) { int t = s[0]; s[0] = s[1]; s[1] = t; }
        return;
    }
    int pivot = s[d_random(m, (unsigned)n)];
    int left = -1, right = n;
    while (left < right) {
        do { left++; } while (left <= right && d_less(m, s[left], pivot, 0));
        do { right--; } while (left <= right && d_less(m, pivot, s[right], 0));
        if (left < right) { int t = s[left]; s[left] = s[right]; s[right] = t; }
    }
    if (left > 1) d_sort(m, s, left);
    if (right < n - 2) d_sort(m, s + right + 1, n - right - 1);
}

/* vertexmedian, triangle.cpp:5513-5569: randomised selection on the chosen axis */
static void d_median(dmesh* m, int* s, int n, int median, int axis)
{
    if (n == 2) {
        if (d_less(m, s[1], s[0], axis)) { int t = s[0]; s[0] = s[1]; s[1] = t; }
        return;
    }
    int pivot = s[d_random(m, (unsigned)n)];
    int left = -1, right = n;
    while (left < right) {
        do { left++; } while (left <= right && d_less(m, s[left], pivot, axis));
        do { right--; } while (left <= right && d_less(m, pivot, s[right], axis));
        if (left < right) { int t = s[left]; s[left] = s[right]; s[right] = t; }
    }
    if (left > median) d_median(m, s, left, median, axis);
    if (right < median - 1) d_median(m, s + right + 1, n - right - 1, median - right - 1, axis);
}

/* alternateaxes, triangle.cpp:5582-5601 */
static void d_alternate(dmesh* m, int* s, int n, int axis)
{
    int divider = n >> 1;
    if (n <= 3) axis = 0;
    d_median(m, s, n, divider, axis);
    if (n - divider >= 2) {
        if (divider >= 2) d_alternate(m, s, divider, 1 - axis);
        d_alternate(m, s + divider, n - divider, 1 - axis);
    }
}

/* mergehulls, triangle.cpp:5638-5934 */
static void d_merge(dmesh* m, otri* farleft, otri* innerleft, otri* innerright, otri* farright, int axis)
{
    otri leftcand, rightcand, baseedge, nextedge, sidecasing, topcasing, outercasing, checkedge;
    int innerleftdest = d_dest(m, *innerleft), innerleftapex = d_apex(m, *innerleft);
    int innerrightorg = d_org(m, *innerright), innerrightapex = d_apex(m, *innerright);
    int farleftpt, farleftapex, farrightpt, farrightapex, checkvertex;
    int lowerleft, lowerright, upperleft, upperright, nextapex;

    if (axis == 1) {                     /* horizontal cut: aim handles at bottom/top-most, :5666-5704 */
        farleftpt = d_org(m, *farleft);   farleftapex = d_apex(m, *farleft);
        farrightpt = d_dest(m, *farright); farrightapex = d_apex(m, *farright);
        while (m->y[farleftapex] < m->y[farleftpt]) {
            *farleft = d_sym(m, d_lnext(*farleft));
            farleftpt = farleftapex;
            farleftapex = d_apex(m, *farleft);
        }
        checkedge = d_sym(m, *innerleft);
        checkvertex = d_apex(m, checkedge);
        while (m->y[checkvertex] > m->y[innerleftdest]) {
            *innerleft = d_lnext(checkedge);
            innerleftapex = innerleftdest;
            innerleftdest = checkvertex;
            checkedge = d_sym(m, *innerleft);
            checkvertex = d_apex(m, checkedge);
        }
        while (m->y[innerrightapex] < m->y[innerrightorg]) {
            *innerright = d_sym(m, d_lnext(*innerright));
            innerrightorg = innerrightapex;
            innerrightapex = d_apex(m, *innerright);
        }
        checkedge = d_sym(m, *farright);
        checkvertex = d_apex(m, checkedge);
        while (m->y[checkvertex] > m->y[farrightpt]) {
            *farright = d_lnext(checkedge);
            farrightapex = farrightpt;
            farrightpt = checkvertex;
            checkedge = d_sym(m, *farright);
            checkvertex = d_apex(m, checkedge);
        }
    }
    int changed;                          /* lower common tangent, :5706-5726 */
    do {
        changed = 0;
        if (d_ccw(m, innerleftdest, innerleftapex, innerrightorg) > 0) {
            *innerleft = d_sym(m, d_lprev(*innerleft));
            innerleftdest = innerleftapex;
            innerleftapex = d_apex(m, *innerleft);
            changed = 1;
        }
        if (d_ccw(m, innerrightapex, innerrightorg, innerleftdest) > 0) {
            *innerright = d_sym(m, d_lnext(*innerright));
            innerrightorg = innerrightapex;
            innerrightapex = d_apex(m, *innerright);
            changed = 1;
        }
    } while (changed);
    leftcand = d_sym(m, *innerleft);      /* :5728-5738 */
    rightcand = d_sym(m, *innerright);
    baseedge = d_make(m);
    d_bond(m, baseedge, *innerleft);
    baseedge = d_lnext(baseedge);
    d_bond(m, baseedge, *innerright);
    baseedge = d_lnext(baseedge);
    d_setorg(m, baseedge, innerrightorg);
    d_setdest(m, baseedge, innerleftdest);
    farleftpt = d_org(m, *farleft);       /* :5745-5752 */
    if (innerleftdest == farleftpt) *farleft = d_lnext(baseedge);
    farrightpt = d_dest(m, *farright);
    if (innerrightorg == farrightpt) *farright = d_lprev(baseedge);
    lowerleft = innerleftdest; lowerright = innerrightorg;
    upperleft = d_apex(m, leftcand); upperright = d_apex(m, rightcand);

    for (;;) {                            /* knit upwards, :5760-5933 */
        int leftfinished = d_ccw(m, upperleft, lowerleft, lowerright) <= 0;
        int rightfinished = d_ccw(m, upperright, lowerleft, lowerright) <= 0;
        if (leftfinished && rightfinished) {
            nextedge = d_make(m);         /* top bounding triangle, :5771-5780 */
            d_setorg(m, nextedge, lowerleft);
            d_setdest(m, nextedge, lowerright);
            d_bond(m, nextedge, baseedge);
            nextedge = d_lnext(nextedge);
            d_bond(m, nextedge, rightcand);
            nextedge = d_lnext(nextedge);
            d_bond(m, nextedge, leftcand);
            if (axis == 1) {              /* restore left/right-most handles, :5786-5809 */
                farleftpt = d_org(m, *farleft);   farleftapex = d_apex(m, *farleft);
                farrightpt = d_dest(m, *farright); farrightapex = d_apex(m, *farright);
                checkedge = d_sym(m, *farleft);
                checkvertex = d_apex(m, checkedge);
                while (m->x[checkvertex] < m->x[farleftpt]) {
                    *farleft = d_lprev(checkedge);
                    farleftapex = farleftpt;
                    farleftpt = checkvertex;
                    checkedge = d_sym(m, *farleft);
                    checkvertex = d_apex(m, checkedge);
                }
                while (m->x[farrightapex] > m->x[farrightpt]) {
                    *farright = d_sym(m, d_lprev(*farright));
                    farrightpt = farrightapex;
                    farrightapex = d_apex(m, *farright);
                }
            }
            return;
        }
        if (!leftfinished) {              /* delete non-Delaunay edges of the left hull, :5813-5859 */
            nextedge = d_sym(m, d_lprev(leftcand));
            nextapex = d_apex(m, nextedge);
            if (nextapex >= 0) {
                int bad = d_incircle(m, lowerleft, lowerright, upperleft, nextapex) > 0;
                while (bad) {
                    nextedge = d_lnext(nextedge);
                    topcasing = d_sym(m, nextedge);
                    nextedge = d_lnext(nextedge);
                    sidecasing = d_sym(m, nextedge);
                    d_bond(m, nextedge, topcasing);
                    d_bond(m, leftcand, sidecasing);
                    leftcand = d_lnext(leftcand);
                    outercasing = d_sym(m, leftcand);
                    nextedge = d_lprev(nextedge);
                    d_bond(m, nextedge, outercasing);
                    d_setorg(m, leftcand, lowerleft);
                    d_setdest(m, leftcand, -1);
                    d_setapex(m, leftcand, nextapex);
                    d_setorg(m, nextedge, -1);
                    d_setdest(m, nextedge, upperleft);
                    d_setapex(m, nextedge, nextapex);
                    upperleft = nextapex;
                    nextedge = sidecasing;
                    nextapex = d_apex(m, nextedge);
                    bad = nextapex >= 0 && d_incircle(m, lowerleft, lowerright, upperleft, nextapex) > 0;
                }
            }
        }
        if (!rightfinished) {             /* and of the right hull, :5861-5907 */
            nextedge = d_sym(m, d_lnext(rightcand));
            nextapex = d_apex(m, nextedge);
            if (nextapex >= 0) {
                int bad = d_incircle(m, lowerleft, lowerright, upperright, nextapex) > 0;
                while (bad) {
                    nextedge = d_lprev(nextedge);
                    topcasing = d_sym(m, nextedge);
                    nextedge = d_lprev(nextedge);
                    sidecasing = d_sym(m, nextedge);
                    d_bond(m, nextedge, topcasing);
                    d_bond(m, rightcand, sidecasing);
                    rightcand = d_lprev(rightcand);
                    outercasing = d_sym(m, rightcand);
                    nextedge = d_lnext(nextedge);
                    d_bond(m, nextedge, outercasing);
                    d_setorg(m, rightcand, -1);
                    d_setdest(m, rightcand, lowerright);
                    d_setapex(m, rightcand, nextapex);
                    d_setorg(m, nextedge, upperright);
                    d_setdest(m, nextedge, -1);
                    d_setapex(m, nextedge, nextapex);
                    upperright = nextapex;
                    nextedge = sidecasing;
                    nextapex = d_apex(m, nextedge);
                    bad = nextapex >= 0 && d_incircle(m, lowerleft, lowerright, upperright, nextapex) > 0;
                }
            }
        }
        if (leftfinished || (!rightfinished &&
                             d_incircle(m, upperleft, lowerleft, lowerright, upperright) > 0)) {
            d_bond(m, baseedge, rightcand);        /* lowerleft -- upperright, :5911-5918 */
            baseedge = d_lprev(rightcand);
            d_setdest(m, baseedge, lowerleft);
            lowerright = upperright;
            rightcand = d_sym(m, baseedge);
            upperright = d_apex(m, rightcand);
        } else {
            d_bond(m, baseedge, leftcand);         /* upperleft -- lowerright, :5920-5927 */
            baseedge = d_lnext(leftcand);
            d_setorg(m, baseedge, lowerright);
            lowerleft = upperleft;
            leftcand = d_sym(m, baseedge);
            upperleft = d_apex(m, leftcand);
        }
    }
}

/* divconqrecurse, triangle.cpp:5953-6103 */
static void d_recurse(dmesh* m, int* s, int n, int axis, otri* farleft, otri* farright)
{
    if (n == 2) {                                   /* :5965-5991 */
        *farleft = d_make(m);
        d_setorg(m, *farleft, s[0]); d_setdest(m, *farleft, s[1]);
        *farright = d_make(m);
        d_setorg(m, *farright, s[1]); d_setdest(m, *farright, s[0]);
        d_bond(m, *farleft, *farright);
        *farleft = d_lprev(*farleft); *farright = d_lnext(*farright);
        d_bond(m, *farleft, *farright);
        *farleft = d_lprev(*farleft); *farright = d_lnext(*farright);
        d_bond(m, *farleft, *farright);
        *farleft = d_lprev(*farright);
        return;
    }
    if (n == 3) {                                   /* :5992-6088 */
        otri midtri = d_make(m), tri1 = d_make(m), tri2 = d_make(m), tri3 = d_make(m);
        int area = d_ccw(m, s[0], s[1], s[2]);
        if (area == 0) {                            /* collinear: two edges */
            d_setorg(m, midtri, s[0]); d_setdest(m, midtri, s[1]);
            d_setorg(m, tri1, s[1]);   d_setdest(m, tri1, s[0]);
            d_setorg(m, tri2, s[2]);   d_setdest(m, tri2, s[1]);
            d_setorg(m, tri3, s[1]);   d_setdest(m, tri3, s[2]);
            d_bond(m, midtri, tri1); d_bond(m, tri2, tri3);
            midtri = d_lnext(midtri); tri1 = d_lprev(tri1); tri2 = d_lnext(tri2); tri3 = d_lprev(tri3);
            d_bond(m, midtri, tri3); d_bond(m, tri1, tri2);
            midtri = d_lnext(midtri); tri1 = d_lprev(tri1); tri2 = d_lnext(tri2); tri3 = d_lprev(tri3);
            d_bond(m, midtri, tri1); d_bond(m, tri2, tri3);
            *farleft = tri1;
            *farright = tri2;
        } else {                                    /* one real triangle + three ghosts */
            d_setorg(m, midtri, s[0]); d_setdest(m, tri1, s[0]); d_setorg(m, tri3, s[0]);
            int p = area > 0 ? s[1] : s[2], q = area > 0 ? s[2] : s[1];
            d_setdest(m, midtri, p); d_setorg(m, tri1, p); d_setdest(m, tri2, p);
            d_setapex(m, midtri, q); d_setorg(m, tri2, q); d_setdest(m, tri3, q);
            d_bond(m, midtri, tri1);
            midtri = d_lnext(midtri);
            d_bond(m, midtri, tri2);
            midtri = d_lnext(midtri);
            d_bond(m, midtri, tri3);
            tri1 = d_lprev(tri1); tri2 = d_lnext(tri2);
            d_bond(m, tri1, tri2);
            tri1 = d_lprev(tri1); tri3 = d_lprev(tri3);
            d_bond(m, tri1, tri3);
            tri2 = d_lnext(tri2); tri3 = d_lprev(tri3);
            d_bond(m, tri2, tri3);
            *farleft = tri1;
            *farright = area > 0 ? tri2 : d_lnext(*farleft);
        }
        return;
    }
    int divider = n >> 1;                           /* :6089-6102 */
    otri innerleft, innerright;
    d_recurse(m, s, divider, 1 - axis, farleft, &innerleft);
    d_recurse(m, s + divider, n - divider, 1 - axis, &innerright, farright);
    d_merge(m, farleft, &innerleft, &innerright, farright, axis);
}

/*
 * computeDelaunayTriangulation (elas.cpp:534-600) -> triangulate("zQB") -> divconqdelaunay
 * (triangle.cpp:6160-6217) -> writeelements (:7800-7853).  sup = n triples (u,v,d); the right
 * image triangulates (u-d, v).  Writes up to cap triangles, returns the triangle count.
 */
int32_t oracle_delaunay(const int32_t* sup, int32_t n, int32_t right_image, int32_t* tri_out, int32_t cap)
{
    if (n < 3) return 0;            /* never reached from process(): elas.cpp:69-75 */
    int32_t* x = (int32_t*)malloc(sizeof(int32_t) * n);
    int32_t* y = (int32_t*)malloc(sizeof(int32_t) * n);
    int* s = (int*)malloc(sizeof(int) * n);
    for (int i = 0; i < n; i++) {
        x[i] = right_image ? sup[3 * i] - sup[3 * i + 2] : sup[3 * i];
        y[i] = sup[3 * i + 1];
        s[i] = i;
    }
    dmesh m;
    m.x = x; m.y = y; m.seed = 1; m.ntri = 0;
    m.nbr = (int*)malloc(sizeof(int) * 3 * (size_t)(3 * n + 8));
    m.vtx = (int*)malloc(sizeof(int) * 3 * (size_t)(3 * n + 8));

    d_sort(&m, s, n);                               /* :6178 */
    int i = 0;                                      /* drop duplicates, keep the first, :6180-6196 */
    for (int j = 1; j < n; j++)
        if (!(x[s[i]] == x[s[j]] && y[s[i]] == y[s[j]])) s[++i] = s[j];
    i++;
    int divider = i >> 1;                           /* :6197-6206 */
    if (i - divider >= 2) {
        if (divider >= 2) d_alternate(&m, s, divider, 1);
        d_alternate(&m, s + divider, i - divider, 1);
    }
    int count = 0;
    if (i >= 2) {
        otri hullleft, hullright;
        d_recurse(&m, s, i, 0, &hullleft, &hullright);   /* :6213 */
        /* removeghosts (:6105-6148) frees exactly the triangles that own the ghost vertex;
         * writeelements then walks the pool in allocation order: corners = (org, dest, apex)
         * at orientation 0. */
        for (int t = 0; t < m.ntri; t++) {
            int a = m.vtx[3 * t + 1], b = m.vtx[3 * t + 2], c = m.vtx[3 * t];
            if (a < 0 || b < 0 || c < 0) continue;
            if (count < cap) { tri_out[3 * count] = a; tri_out[3 * count + 1] = b; tri_out[3 * count + 2] = c; }
            count++;
        }
    }
    free(m.nbr); free(m.vtx); free(s); free(x); free(y);
    return count;
}

/* ------------------------------------------------------------------------------------------ */
/* a11: plane fit.  Matrix::solve (matrix.cpp:414-502) = Gauss-Jordan with full pivoting in    */
/* double, for one 3x3 system and one right-hand side; returns 0 when a pivot is < 1e-20.      */
/* ------------------------------------------------------------------------------------------ */

static int solve3(double A[3][3], double b[3])
{
    int ipiv[3] = {0, 0, 0};
    for (int i = 0; i < 3; i++) {
        double big = 0.0;
        int irow = 0, icol = 0;
        for (int j = 0; j < 3; j++)
            if (ipiv[j] != 1)
                for (int k = 0; k < 3; k++)
                    if (ipiv[k] == 0 && fabs(A[j][k]) >= big) { big = fabs(A[j][k]); irow = j; icol = k; }
        ++ipiv[icol];
        if (irow != icol) {
            for (int l = 0; l < 3; l++) { double t = A[irow][l]; A[irow][l] = A[icol][l]; A[icol][l] = t; }
            double t = b[irow]; b[irow] = b[icol]; b[icol] = t;
        }
        if (fabs(A[icol][icol]) < 1e-20) return 0;
        double pivinv = 1.0 / A[icol][icol];
        A[icol][icol] = 1.0;
        for (int l = 0; l < 3; l++) A[icol][l] *= pivinv;
        b[icol] *= pivinv;
        for (int ll = 0; ll < 3; ll++)
            if (ll != icol) {
                double dum = A[ll][icol];
                A[ll][icol] = 0.0;
                for (int l = 0; l < 3; l++) A[ll][l] -= A[icol][l] * dum;
                b[ll] -= b[icol] * dum;
            }
    }
    return 1;      /* the column unscrambling (:489-495) touches only A */
}

/* elas.cpp:605-680: planes[6*i..] = t1a,t1b,t1c (left coords), t2a,t2b,t2c (right coords) */
static void disparity_planes(const int32_t* sup, const int32_t* tri, int nt, float* planes)
{
    for (int i = 0; i < nt; i++)
        for (int k = 0; k < 2; k++) {
            double A[3][3], b[3];
            for (int c = 0; c < 3; c++) {
                const int32_t* s = sup + 3 * tri[3 * i + c];
                A[c][0] = k ? s[0] - s[2] : s[0];
                A[c][1] = s[1];
                A[c][2] = 1;
                b[c] = s[2];
            }
            float* o = planes + 6 * i + 3 * k;
            if (solve3(A, b)) { o[0] = (float)b[0]; o[1] = (float)b[1]; o[2] = (float)b[2]; }
            else o[0] = o[1] = o[2] = 0;
        }
}

/* ------------------------------------------------------------------------------------------ */
/* a12: candidate grid (elas.cpp:684-780), including the flat-index diffusion     */
/* ------------------------------------------------------------------------------------------ */

static void create_grid(const elas_b200_params* p, const int32_t* sup, int n, int32_t* grid,
                        int gw, int gh, int right_image)
{
    int dn = p->disp_max + 1;
    size_t cells = (size_t)gw * gh;
    int32_t* t1 = (int32_t*)calloc(cells * dn, sizeof(int32_t));
    int32_t* t2 = (int32_t*)calloc(cells * dn, sizeof(int32_t));
    for (int i = 0; i < n; i++) {
        int xc = sup[3 * i], yc = sup[3 * i + 1], dc = sup[3 * i + 2];
        int x = right_image ? (int)floorf((float)(xc - dc) / (float)p->grid_size)
                            : (int)floorf((float)(xc / p->grid_size));          /* :712,:716 */
        int y = (int)floorf((float)yc / (float)p->grid_size);
        if (x < 0 || x >= gw || y < 0 || y >= gh) continue;
        for (int d = MAXI(dc - 1, 0); d <= MINI(dc + 1, p->disp_max); d++) t1[((size_t)y * gw + x) * dn + d] = 1;
    }
    /* :732-751: nine pointers walk temp1 in lock-step; out cell c (flat) = OR of cells
     * c + {-gw-1,-gw,-gw+1,-1,0,1,gw-1,gw,gw+1}, for c in [gw+1, gw*gh-gw-2]. */
    for (long c = gw + 1; c <= (long)cells - gw - 2; c++) {
        static const int oy[9] = {-1, -1, -1, 0, 0, 0, 1, 1, 1}, ox[9] = {-1, 0, 1, -1, 0, 1, -1, 0, 1};
        for (int d = 0; d < dn; d++) {
            int32_t v = 0;
            for (int k = 0; k < 9; k++) v |= t1[(size_t)(c + oy[k] * gw + ox[k]) * dn + d];
            t2[(size_t)c * dn + d] = v;
        }
    }
    memset(grid, 0, cells * (dn + 1) * sizeof(int32_t));
    for (size_t c = 0; c < cells; c++) {                 /* :754-775 */
        int k = 1;
        for (int d = 0; d < dn; d++)
            if (t2[c * dn + d] > 0) grid[c * (dn + 1) + k++] = d;
        grid[c * (dn + 1)] = k - 1;
    }
    free(t1); free(t2);
}

/* ------------------------------------------------------------------------------------------ */
/* a13 + a14: dense matching (elas.cpp:960-1118, findMatch :814-955)                           */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    const elas_b200_params* p;
    int W, H, gw, dn;                /* dn = disp_max + 1 (= grid_dims[0]-1, :819) */
    const uint8_t *own, *other;
    const int32_t* grid;
    const int32_t* P;
    int plane_radius, right_image;
    float* D;
} match_ctx;

static void find_match(const match_ctx* c, int u, int v, float pa, float pb, float pc, int valid)
{
    const elas_b200_params* p = c->p;
    const int W = c->W, H = c->H, window = 2;
    size_t d_addr = p->subsampling ? (size_t)(v / 2) * (W / 2) + u / 2 : (size_t)v * W + u;   /* :824-825 */
    if (u < window || u >= W - window) return;                                               /* :828 */
    size_t line = (size_t)16 * W * MAXI(MINI(v, H - 3), 2);                                   /* :834 */
    const uint8_t* own = c->own + line + 16 * (size_t)u;
    const uint8_t* other_line = c->other + line;
    if (texture16(own) < p->match_texture) return;                                           /* :851-859 */

    int d_plane = (int)(pa * (float)u + pb * (float)v + pc);                                  /* :861 */
    int d_plane_min = MAXI(d_plane - c->plane_radius, 0);
    int d_plane_max = MINI(d_plane + c->plane_radius, c->dn - 1);
    int gx = (int)floorf((float)u / (float)p->grid_size);                                     /* :866-867 */
    int gy = (int)floorf((float)v / (float)p->grid_size);
    const int32_t* cell = c->grid + ((size_t)gy * c->gw + gx) * (c->dn + 1);
    int num_grid = cell[0];

    int min_val = 10000, min_d = -1;                                                          /* :878-879 */
    for (int i = 0; i < num_grid; i++) {                                                      /* :890-903 */
        int d = cell[1 + i];
        if (d < d_plane_min || d > d_plane_max) {
            int uw = c->right_image ? u + d : u - d;
            if (uw < window || uw >= W - window) continue;
            int val = sad16(own, other_line + 16 * (size_t)uw);
            if (val < min_val) { min_val = val; min_d = d; }
        }
    }
    for (int d = d_plane_min; d <= d_plane_max; d++) {                                        /* :904-913 */
        int uw = c->right_image ? u + d : u - d;
        if (uw < window || uw >= W - window) continue;
        int val = sad16(own, other_line + 16 * (size_t)uw) + (valid ? c->P[abs(d - d_plane)] : 0);
        if (val < min_val) { min_val = val; min_d = d; }
    }
    c->D[d_addr] = min_d >= 0 ? (float)min_d : -1.0f;                                          /* :947-954 */
}

/* The float expression (uint32_t)(a*u+b) stored to int32 (elas.cpp:1081-1082): gcc on x86-64
 * converts through a 64-bit truncation, i.e. the low 32 bits of (int64)trunc(x). */
static int32_t trunc_u32(float x) { return (int32_t)(uint32_t)(int64_t)x; }

static void compute_disparity(const elas_b200_params* p, int W, int H, const int32_t* sup,
                              const int32_t* tri, const float* planes, int nt,
                              const int32_t* grid, int gw, const uint8_t* desc1, const uint8_t* desc2,
                              int right_image, float* D)
{
    const int dn = p->disp_max + 1;
    size_t nd = p->subsampling ? (size_t)(W / 2) * (H / 2) : (size_t)W * H;
    for (size_t i = 0; i < nd; i++) D[i] = -10;                                               /* :968-981 */

    float two_sigma_squared = 2 * p->sigma * p->sigma;                                        /* :984-992 */
    int32_t* P = (int32_t*)malloc(sizeof(int32_t) * dn);
    for (int dd = 0; dd < dn; dd++) {
        float tmp = -logf(p->gamma + expf(-dd * dd / two_sigma_squared)) + logf(p->gamma);
        P[dd] = (int32_t)(tmp / p->beta);
    }
    int plane_radius = (int)fmaxf(ceilf(p->sigma * p->sradius), 2.0f);                        /* :993 */

    match_ctx c = {p, W, H, gw, dn, right_image ? desc2 : desc1, right_image ? desc1 : desc2,
                   grid, P, plane_radius, right_image, D};

    for (int i = 0; i < nt; i++) {                                                            /* :1003 */
        const float* pl = planes + 6 * i;
        float pa = right_image ? pl[3] : pl[0], pb = right_image ? pl[4] : pl[1];
        float pc = right_image ? pl[5] : pl[2], pd = right_image ? pl[0] : pl[3];
        float tu[3], tv[3];
        for (int k = 0; k < 3; k++) {
            const int32_t* s = sup + 3 * tri[3 * i + k];
            tu[k] = right_image ? (float)(s[0] - s[2]) : (float)s[0];
            tv[k] = (float)s[1];
        }
        for (int j = 0; j < 3; j++)                                                           /* :1043-1053 */
            for (int k = 0; k < j; k++)
                if (tu[k] > tu[j]) {
                    float t = tu[j]; tu[j] = tu[k]; tu[k] = t;
                    t = tv[j]; tv[j] = tv[k]; tv[k] = t;
                }
        float Au = tu[0], Av = tv[0], Bu = tu[1], Bv = tv[1], Cu = tu[2], Cv = tv[2];
        float ABa = 0, ACa = 0, BCa = 0;                                                      /* :1061-1067 */
        if ((int)Au != (int)Bu) ABa = (Av - Bv) / (Au - Bu);
        if ((int)Au != (int)Cu) ACa = (Av - Cv) / (Au - Cu);
        if ((int)Bu != (int)Cu) BCa = (Bv - Cv) / (Bu - Cu);
        float ABb = Av - ABa * Au, ACb = Av - ACa * Au, BCb = Bv - BCa * Bu;
        int valid = fabs(pa) < 0.7 && fabs(pd) < 0.7;                                         /* :1072 */

        for (int half = 0; half < 2; half++) {                                                /* :1074-1114 */
            float u0 = half ? Bu : Au, u1 = half ? Cu : Bu;
            float ea = half ? BCa : ABa, eb = half ? BCb : ABb;
            if ((int)u0 == (int)u1) continue;
            for (int u = MAXI((int)u0, 0); u < MINI((int)u1, W); u++) {
                if (p->subsampling && u % 2) continue;
                int v1 = trunc_u32(ACa * (float)u + ACb);
                int v2 = trunc_u32(ea * (float)u + eb);
                for (int v = MINI(v1, v2); v < MAXI(v1, v2); v++) {
                    if (p->subsampling && v % 2) continue;
                    if (v < 0 || v >= H) continue;     /* never taken for hull-interior triangles */
                    find_match(&c, u, v, pa, pb, pc, valid);
                }
            }
        }
    }
    free(P);
}

/* ------------------------------------------------------------------------------------------ */
/* a15: left/right consistency (elas.cpp:1122-1204)                                            */
/* ------------------------------------------------------------------------------------------ */

static void lr_check(const elas_b200_params* p, int Dw, int Dh, float* D1, float* D2)
{
    size_t n = (size_t)Dw * Dh;
    float* C1 = (float*)malloc(n * sizeof(float));
    float* C2 = (float*)malloc(n * sizeof(float));
    memcpy(C1, D1, n * sizeof(float)); memcpy(C2, D2, n * sizeof(float));
    for (int u = 0; u < Dw; u++)
        for (int v = 0; v < Dh; v++) {
            size_t a = (size_t)v * Dw + u;
            float d1 = C1[a], d2 = C2[a];
            float w1 = p->subsampling ? (float)u - d1 / 2 : (float)u - d1;
            float w2 = p->subsampling ? (float)u + d2 / 2 : (float)u + d2;
            if (d1 >= 0 && w1 >= 0 && w1 < Dw) {
                if (fabs(C2[(size_t)v * Dw + (int)w1] - d1) > p->lr_threshold) D1[a] = -10;
            } else D1[a] = -10;
            if (d2 >= 0 && w2 >= 0 && w2 < Dw) {
                if (fabs(C1[(size_t)v * Dw + (int)w2] - d2) > p->lr_threshold) D2[a] = -10;
            } else D2[a] = -10;
        }
    free(C1); free(C2);
}

/* ------------------------------------------------------------------------------------------ */
/* a16: speckle removal (elas.cpp:1208-1326): 4-connected flood fill, edge iff |delta| <= thr   */
/* ------------------------------------------------------------------------------------------ */

static void remove_small_segments(const elas_b200_params* p, int Dw, int Dh, float* D)
{
    int speckle = p->speckle_size;
    if (p->subsampling) speckle = (int)(sqrt((float)p->speckle_size) * 2);    /* :1218 */
    size_t n = (size_t)Dw * Dh;
    int32_t* done = (int32_t*)calloc(n, sizeof(int32_t));
    int32_t* lu = (int32_t*)calloc(n, sizeof(int32_t));
    int32_t* lv = (int32_t*)calloc(n, sizeof(int32_t));
    for (int u = 0; u < Dw; u++)
        for (int v = 0; v < Dh; v++) {
            if (done[(size_t)v * Dw + u]) continue;
            lu[0] = u; lv[0] = v;
            int count = 1, curr = 0;
            while (curr < count) {
                int uc = lu[curr], vc = lv[curr];
                size_t ac = (size_t)vc * Dw + uc;
                const int nu[4] = {uc - 1, uc + 1, uc, uc}, nv[4] = {vc, vc, vc - 1, vc + 1};
                for (int i = 0; i < 4; i++) {
                    if (nu[i] < 0 || nv[i] < 0 || nu[i] >= Dw || nv[i] >= Dh) continue;
                    size_t an = (size_t)nv[i] * Dw + nu[i];
                    if (done[an] == 0 && D[an] >= 0 && fabs(D[ac] - D[an]) <= p->speckle_sim_threshold) {
                        lu[count] = nu[i]; lv[count] = nv[i]; count++;
                        done[an] = 1;
                    }
                }
                curr++;
                done[ac] = 1;
            }
            if (count < speckle)
                for (int i = 0; i < count; i++) D[(size_t)lv[i] * Dw + lu[i]] = -10;
        }
    free(done); free(lu); free(lv);
}

/* ------------------------------------------------------------------------------------------ */
/* a17: gap interpolation (elas.cpp:1330-1530)                                                 */
/* ------------------------------------------------------------------------------------------ */

static void gap_line(const elas_b200_params* p, float* D, int len, ptrdiff_t stride, int gap)
{
    const float discon = 3.0f;
    int count = 0;
    for (int i = 0; i < len; i++) {
        if (D[i * stride] >= 0) {
            if (count >= 1 && count <= gap) {
                int first = i - count, last = i - 1;
                if (first > 0 && last < len - 1) {
                    float d1 = D[(first - 1) * stride], d2 = D[(last + 1) * stride];
                    float dip = fabs(d1 - d2) < discon ? (d1 + d2) / 2 : (d1 < d2 ? d1 : d2);
                    for (int k = first; k <= last; k++) D[k * stride] = dip;
                }
            }
            count = 0;
        } else count++;
    }
    if (p->add_corners) {                              /* :1401-1436, :1493-1528 */
        for (int i = 0; i < len; i++)
            if (D[i * stride] >= 0) {
                for (int k = MAXI(i - gap, 0); k < i; k++) D[k * stride] = D[i * stride];
                break;
            }
        for (int i = len - 1; i >= 0; i--)
            if (D[i * stride] >= 0) {
                for (int k = i; k <= MINI(i + gap, len - 1); k++) D[k * stride] = D[i * stride];
                break;
            }
    }
}

static void gap_interpolation(const elas_b200_params* p, int Dw, int Dh, float* D)
{
    int gap = p->subsampling ? p->ipol_gap_width / 2 + 1 : p->ipol_gap_width;   /* :1335-1341 */
    for (int v = 0; v < Dh; v++) gap_line(p, D + (size_t)v * Dw, Dw, 1, gap);
    for (int u = 0; u < Dw; u++) gap_line(p, D + u, Dh, Dw, gap);
}

/* ------------------------------------------------------------------------------------------ */
/* a18: "adaptive mean" (elas.cpp:1535-1754).  The reference's abs mask is the float           */
/* 2147483648.0f = bits 0x4F000000, so the weight max(0, 4 - M(x - x_c)) takes the  */
/* values 4, 2, 0.  Sums follow the SSE lane order: lane k = slot k + slot k+4, then            */
/* ((l0+l1)+l2)+l3.                                                                             */
/* ------------------------------------------------------------------------------------------ */

static float masked(float x)
{
    uint32_t b;
    memcpy(&b, &x, 4);
    b &= 0x4F000000u;
    memcpy(&x, &b, 4);
    return x;
}

static void mean_taps(const float* val, int taps, float centre, float* wsum, float* fsum)
{
    float w[8], f[8];
    for (int k = 0; k < taps; k++) {
        float m = 4.0f - masked(val[k] - centre);
        w[k] = m > 0.0f ? m : 0.0f;              /* _mm_max_ps(0, x) */
        f[k] = val[k] * w[k];
    }
    if (taps == 8) {
        float lw[4], lf[4];
        for (int k = 0; k < 4; k++) { lw[k] = w[k] + w[k + 4]; lf[k] = f[k] + f[k + 4]; }
        *wsum = lw[0] + lw[1] + lw[2] + lw[3];
        *fsum = lf[0] + lf[1] + lf[2] + lf[3];
    } else {
        *wsum = w[0] + w[1] + w[2] + w[3];
        *fsum = f[0] + f[1] + f[2] + f[3];
    }
}

static void adaptive_mean(const elas_b200_params* p, int Dw, int Dh, float* D)
{
    size_t n = (size_t)Dw * Dh;
    float* C = (float*)malloc(n * sizeof(float));
    float* T = (float*)malloc(n * sizeof(float));
    memcpy(C, D, n * sizeof(float));
    memcpy(T, D, n * sizeof(float));     /* reference: malloc'ed, set only where D<0 (A.9); defined here */
    for (size_t i = 0; i < n; i++) if (D[i] < 0) { C[i] = -10; T[i] = -10; }
    const int taps = p->subsampling ? 4 : 8;       /* :1574 / :1651 */
    const int lag = p->subsampling ? 1 : 3;        /* centre = newest - lag */
    float val[8];
    for (int v = 3; v < Dh - 3; v++) {             /* horizontal, :1654-1698 (:1577-1611) */
        for (int u = 0; u < taps - 1; u++) val[u] = C[(size_t)v * Dw + u];
        for (int u = taps - 1; u < Dw; u++) {
            float centre = C[(size_t)v * Dw + (u - lag)];
            val[u % taps] = C[(size_t)v * Dw + u];
            float ws, fs;
            mean_taps(val, taps, centre, &ws, &fs);
            if (ws > 0) { float d = fs / ws; if (d >= 0) T[(size_t)v * Dw + (u - lag)] = d; }
        }
    }
    for (int u = 3; u < Dw - 3; u++) {             /* vertical, :1701-1745 (:1614-1648) */
        for (int v = 0; v < taps - 1; v++) val[v] = T[(size_t)v * Dw + u];
        for (int v = taps - 1; v < Dh; v++) {
            float centre = T[(size_t)(v - lag) * Dw + u];
            val[v % taps] = T[(size_t)v * Dw + u];
            float ws, fs;
            mean_taps(val, taps, centre, &ws, &fs);
            if (ws > 0) { float d = fs / ws; if (d >= 0) D[(size_t)(v - lag) * Dw + u] = d; }
        }
    }
    free(C); free(T);
}

/* ------------------------------------------------------------------------------------------ */
/* a19: separable 7-tap median (elas.cpp:1758-1838)                                            */
/* ------------------------------------------------------------------------------------------ */

static float median7(const float* src, ptrdiff_t stride)
{
    float vals[7];
    for (int j = 0; j < 7; j++) {
        float t = src[(j - 3) * stride];
        int i = j - 1;
        while (i >= 0 && vals[i] > t) { vals[i + 1] = vals[i]; i--; }
        vals[i + 1] = t;
    }
    return vals[3];
}

static void median_filter(int Dw, int Dh, float* D)
{
    float* T = (float*)calloc((size_t)Dw * Dh, sizeof(float));
    for (int u = 3; u < Dw - 3; u++)
        for (int v = 3; v < Dh - 3; v++) {
            size_t a = (size_t)v * Dw + u;
            T[a] = D[a] >= 0 ? median7(D + a, 1) : D[a];
        }
    for (int u = 3; u < Dw - 3; u++)
        for (int v = 3; v < Dh - 3; v++) {
            size_t a = (size_t)v * Dw + u;
            if (D[a] >= 0) D[a] = median7(T + a, Dw);
        }
    free(T);
}

/* ------------------------------------------------------------------------------------------ */
/* Elas::process (elas.cpp:32-170)                                                             */
/* ------------------------------------------------------------------------------------------ */

static int32_t run(const elas_b200_params* p, const uint8_t* I1_, const uint8_t* I2_,
                   float* D1, float* D2, const int32_t* dims)
{
    const int W = dims[0], H = dims[1];
    const int bpl = W + 15 - (W - 1) % 16;                                    /* :37 */
    uint8_t* I1 = (uint8_t*)calloc((size_t)bpl * H, 1);
    uint8_t* I2 = (uint8_t*)calloc((size_t)bpl * H, 1);
    if (bpl == dims[2]) { memcpy(I1, I1_, (size_t)bpl * H); memcpy(I2, I2_, (size_t)bpl * H); }   /* :44-48 */
    else for (int v = 0; v < H; v++) {                                        /* :49-56 */
        memcpy(I1 + (size_t)v * bpl, I1_ + (size_t)v * dims[2], W);
        memcpy(I2 + (size_t)v * bpl, I2_ + (size_t)v * dims[2], W);
    }
    uint8_t* desc1 = (uint8_t*)malloc((size_t)16 * W * H);
    uint8_t* desc2 = (uint8_t*)malloc((size_t)16 * W * H);
    descriptor(I1, W, H, bpl, p->subsampling, desc1);                         /* :61-62 */
    descriptor(I2, W, H, bpl, p->subsampling, desc2);
    keep("desc1", desc1, (int64_t)16 * W * H);
    keep("desc2", desc2, (int64_t)16 * W * H);
    free(I1); free(I2);

    int step = p->candidate_stepsize + (p->subsampling ? p->candidate_stepsize % 2 : 0);
    size_t cap = ((size_t)(W / step) + 2) * ((size_t)(H / step) + 2) + 6;
    int32_t* sup = (int32_t*)malloc(cap * 3 * sizeof(int32_t));
    int Wc, Hc;
    int n = support_matches(p, W, H, desc1, desc2, sup, &Wc, &Hc);             /* :66 */
    keep("support", sup, (int64_t)n * 12);
    int32_t rc = 0;
    if (n < 3) { rc = ELAS_B200_E_FEW_SUPPORT; goto done; }                    /* :69-75 */

    {
        int cap_t = 2 * n + 8;
        int32_t* tri1 = (int32_t*)malloc((size_t)cap_t * 12);
        int32_t* tri2 = (int32_t*)malloc((size_t)cap_t * 12);
        int nt1 = oracle_delaunay(sup, n, 0, tri1, cap_t);                     /* :80-81 */
        int nt2 = oracle_delaunay(sup, n, 1, tri2, cap_t);
        float* pl1 = (float*)malloc((size_t)cap_t * 24);
        float* pl2 = (float*)malloc((size_t)cap_t * 24);
        disparity_planes(sup, tri1, nt1, pl1);                                 /* :87-88 */
        disparity_planes(sup, tri2, nt2, pl2);
        keep("tri1", tri1, (int64_t)nt1 * 12); keep("tri2", tri2, (int64_t)nt2 * 12);
        keep("planes1", pl1, (int64_t)nt1 * 24); keep("planes2", pl2, (int64_t)nt2 * 24);

        int gw = (int)ceilf((float)W / (float)p->grid_size);                   /* :98-105 */
        int gh = (int)ceilf((float)H / (float)p->grid_size);
        size_t gsz = (size_t)(p->disp_max + 2) * gw * gh;
        int32_t* g1 = (int32_t*)malloc(gsz * 4);
        int32_t* g2 = (int32_t*)malloc(gsz * 4);
        create_grid(p, sup, n, g1, gw, gh, 0);
        create_grid(p, sup, n, g2, gw, gh, 1);
        int32_t gd[3] = {p->disp_max + 2, gw, gh};
        keep("grid1", g1, (int64_t)gsz * 4); keep("grid2", g2, (int64_t)gsz * 4);
        keep("grid_dims", gd, sizeof gd);

        compute_disparity(p, W, H, sup, tri1, pl1, nt1, g1, gw, desc1, desc2, 0, D1);   /* :110-111 */
        compute_disparity(p, W, H, sup, tri2, pl2, nt2, g2, gw, desc1, desc2, 1, D2);
        int Dw = p->subsampling ? W / 2 : W, Dh = p->subsampling ? H / 2 : H;
        int64_t nb = (int64_t)Dw * Dh * 4;
        keep("D1_raw", D1, nb); keep("D2_raw", D2, nb);

        lr_check(p, Dw, Dh, D1, D2);                                            /* :116 */
        keep("D1_lr", D1, nb); keep("D2_lr", D2, nb);
        remove_small_segments(p, Dw, Dh, D1);                                   /* :121-125 */
        if (!p->postprocess_only_left) remove_small_segments(p, Dw, Dh, D2);
        keep("D1_seg", D1, nb); keep("D2_seg", D2, nb);
        gap_interpolation(p, Dw, Dh, D1);                                       /* :130-134 */
        if (!p->postprocess_only_left) gap_interpolation(p, Dw, Dh, D2);
        keep("D1_gap", D1, nb); keep("D2_gap", D2, nb);
        if (p->filter_adaptive_mean) {                                          /* :136-146 */
            adaptive_mean(p, Dw, Dh, D1);
            if (!p->postprocess_only_left) adaptive_mean(p, Dw, Dh, D2);
        }
        keep("D1_mean", D1, nb); keep("D2_mean", D2, nb);
        if (p->filter_median) {                                                 /* :148-159 */
            median_filter(Dw, Dh, D1);
            if (!p->postprocess_only_left) median_filter(Dw, Dh, D2);
        }
        keep("D1", D1, nb); keep("D2", D2, nb);
        free(tri1); free(tri2); free(pl1); free(pl2); free(g1); free(g2);
    }
done:
    free(sup); free(desc1); free(desc2);
    return rc;
}

int32_t oracle_process(const elas_b200_params* p, const uint8_t* I1, const uint8_t* I2,
                       float* D1, float* D2, const int32_t* dims)
{
    g_keep = 0;
    return run(p, I1, I2, D1, D2, dims);
}

int32_t oracle_run_stages(const elas_b200_params* p, const uint8_t* I1, const uint8_t* I2,
                          float* D1, float* D2, const int32_t* dims)
{
    stage_clear();
    g_keep = 1;
    int32_t rc = run(p, I1, I2, D1, D2, dims);
    g_keep = 0;
    return rc;
}

/* Single stages on caller-provided tables, so a CUDA stage can be checked on reference inputs. */
void oracle_descriptor(const uint8_t* I, int32_t W, int32_t H, int32_t bpl, int32_t half, uint8_t* desc)
{
    descriptor(I, W, H, bpl, half, desc);
}

void oracle_lattice_filters(const elas_b200_params* p, int16_t* dcan, int32_t Wc, int32_t Hc)
{
    remove_inconsistent(p, dcan, Wc, Hc);
    remove_redundant(dcan, Wc, Hc, 5, 1, 1);
    remove_redundant(dcan, Wc, Hc, 5, 1, 0);
}

void oracle_planes(const int32_t* sup, const int32_t* tri, int32_t nt, float* planes)
{
    disparity_planes(sup, tri, nt, planes);
}

/* =============================================================================================
 * the consumers of D1 inside StereoThread (pinned against the reference's own
 * statements through oracle/_ref/libview_ref.so, tests/test_view.py).
 * ============================================================================================= */

/* HSV colour map of min(D1/200, 1), stereothread.cpp:116-147.  out = 3 floats per pixel.
 * Expression types follow the C++ source: h2 and x are doubles narrowed to float, fmod is the float
 * overload (exact either way). */
void oracle_colormap(const float* D1, int32_t d_width, int32_t d_height, float* out)
{
    const float d_max = 200;                                                   /* :117 */
    for (int32_t i = 0; i < d_width * d_height; i++) {
        float* o = out + 3 * (size_t)i;
        float val = D1[i] / d_max;                                             /* :130 */
        if ((float)1.0 < val) val = (float)1.0;                                /* std::min(a,b): b < a ? b : a */
        if (val <= 0) { o[0] = 0; o[1] = 0; o[2] = 0; continue; }              /* :131-134 */
        const float h2 = (float)(6.0 * (1.0 - (double)val));                   /* :137 */
        const float x = (float)(1.0 * (1.0 - fabs((double)fmodf(h2, (float)2.0) - 1.0)));   /* :138 */
        if      (0 <= h2 && h2 < 1)  { o[0] = 1; o[1] = x; o[2] = 0; }         /* :139-144 */
        else if (1 <= h2 && h2 < 2)  { o[0] = x; o[1] = 1; o[2] = 0; }
        else if (2 <= h2 && h2 < 3)  { o[0] = 0; o[1] = 1; o[2] = x; }
        else if (3 <= h2 && h2 < 4)  { o[0] = 0; o[1] = x; o[2] = 1; }
        else if (4 <= h2 && h2 < 5)  { o[0] = x; o[1] = 0; o[2] = 1; }
        else if (5 <= h2 && h2 <= 6) { o[0] = 1; o[1] = 0; o[2] = x; }
        else                         { o[0] = 0; o[1] = 0; o[2] = 0; }         /* unreachable for val in (0,1] */
    }
}

/* Back-projection of D1 and the intensity image with its border gain, StereoThread::createCurrentMap,
 * stereothread.cpp:180-255.  view = {f, cu, cv, base, max_dist, gain}; H = rows 0..2 of the pose
 * (3x4, row-major, double like libviso2's Matrix).  X/Y/Z are 0 where the reference writes nothing. */
void oracle_reproject(const uint8_t* I1, const float* D1, int32_t width, int32_t height, int32_t step,
                      const float* view, const double* H, float* I, float* D, float* X, float* Y, float* Z)
{
    const float f = view[0], cu = view[1], cv = view[2], base = view[3], max_dist = view[4], gain = view[5];
    float h[12];
    for (int k = 0; k < 12; k++) h[k] = (float)H[k];                           /* :201-204 */
    for (int32_t v = 0; v < height; v++)                                       /* :193-200 */
        for (int32_t u = 0; u < width; u++) {
            const size_t a = (size_t)v * width + u;
            D[a] = D1[a];
            I[a] = (float)(((float)I1[(size_t)v * step + u]) / 255.0);
            X[a] = Y[a] = Z[a] = 0.f;
        }
    for (int32_t u = 0; u < width; u++)                                        /* :205-229 */
        for (int32_t v = 0; v < height; v++) {
            const size_t a = (size_t)v * width + u;
            const float d = D[a];
            if (d > 0) {
                const float z = (f * base) / d;
                if (((double)z > 0.1) && (z < max_dist)) {
                    const float x = ((float)u - cu) * base / d;
                    const float y = ((float)v - cv) * base / d;
                    X[a] = h[0] * x + h[1] * y + h[2] * z + h[3];
                    Y[a] = h[4] * x + h[5] * y + h[6] * z + h[7];
                    Z[a] = h[8] * x + h[9] * y + h[10] * z + h[11];
                } else {
                    D[a] = -1;
                }
            }
        }
    /* gain ramp on the image border, :231-252 */
    int32_t margin = 200 < width / 2 ? 200 : width / 2;
    if (height / 2 < margin) margin = height / 2;
    float gain_inv = 1;
    if (gain) gain_inv = (float)(1.0 / gain);
    for (int32_t i = 0; i < margin; i++) {
        const float g = (float)(((float)(margin - i) * gain_inv + (float)i * 1.0) / (float)margin);
        for (int32_t u = margin; u < width - margin; u++) {
            float* p0 = &I[(size_t)i * width + u];
            float* p1 = &I[(size_t)(height - i - 1) * width + u];
            float t = g * *p0; t = t < 0.f ? 0.f : t; *p0 = 1.f < t ? 1.f : t;
            t = g * *p1; t = t < 0.f ? 0.f : t; *p1 = 1.f < t ? 1.f : t;
        }
        for (int32_t v = margin; v < height - margin; v++) {
            float* p0 = &I[(size_t)v * width + i];
            float* p1 = &I[(size_t)v * width + width - i - 1];
            float t = g * *p0; t = t < 0.f ? 0.f : t; *p0 = 1.f < t ? 1.f : t;
            t = g * *p1; t = t < 0.f ? 0.f : t; *p1 = 1.f < t ? 1.f : t;
        }
    }
}

/* Matrix::inv of a 4x4 (libviso2/src/matrix.cpp:593-604 = eye(4).solve(A), :648-757): Gauss-Jordan elimination with
 * full pivoting in double; B = identity on entry, the inverse on exit.  Returns 0 when a pivot is below 1e-20. */
static int oracle_inv4(double A[4][4], double B[4][4])
{
    int ipiv[4] = {0, 0, 0, 0};
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) B[i][j] = i == j;
    int irow = 0, icol = 0;
    for (int i = 0; i < 4; i++) {
        double big = 0.0;
        for (int j = 0; j < 4; j++)
            if (ipiv[j] != 1)
                for (int k = 0; k < 4; k++)
                    if (ipiv[k] == 0 && fabs(A[j][k]) >= big) { big = fabs(A[j][k]); irow = j; icol = k; }
        ++ipiv[icol];
        if (irow != icol)
            for (int l = 0; l < 4; l++) {
                double t = A[irow][l]; A[irow][l] = A[icol][l]; A[icol][l] = t;
                t = B[irow][l]; B[irow][l] = B[icol][l]; B[icol][l] = t;
            }
        if (fabs(A[icol][icol]) < 1e-20) return 0;
        const double pivinv = 1.0 / A[icol][icol];
        A[icol][icol] = 1.0;
        for (int l = 0; l < 4; l++) A[icol][l] *= pivinv;
        for (int l = 0; l < 4; l++) B[icol][l] *= pivinv;
        for (int ll = 0; ll < 4; ll++)
            if (ll != icol) {
                const double dum = A[ll][icol];
                A[ll][icol] = 0.0;
                for (int l = 0; l < 4; l++) A[ll][l] -= A[icol][l] * dum;
                for (int l = 0; l < 4; l++) B[ll][l] -= B[icol][l] * dum;
            }
    }
    return 1;
}

/* Fusion of the current map with the previous one, StereoThread::addDisparityMapToReconstruction,
 * stereothread.cpp:290-437.  The previous map (p*, may all be NULL) is the fused current map of the call before
 * (the reference's own hand-over, :433-434, frees what it has just copied; see oracle/ref_view_harness.cpp); pD comes
 * back with the merged points invalidated.  The current map (c*, from oracle_reproject) is fused in place.
 * view = {f, cu, cv, base, max_dist, gain}; H = rows 0..2 of the current pose.  points_*: (x,y,z,val) quadruples. */
void oracle_fuse(int32_t width, int32_t height, const float* view, const double* H,
                 float* pI, float* pD, float* pX, float* pY, float* pZ,
                 float* cI, float* cD, float* cX, float* cY, float* cZ,
                 float* points_prev, int32_t* n_prev, float* points_curr, int32_t* n_curr)
{
    const float max_dist = view[4];
    *n_prev = 0;
    if (pI && pD && pX && pY && pZ) {                                           /* :296-300 */
        double A[4][4] = {{0}}, Hi[4][4];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 4; c++) A[r][c] = H[4 * r + c];
        A[3][3] = 1.0;
        if (!oracle_inv4(A, Hi)) { /* Matrix::solve returns false and leaves B half reduced; poses are never singular */ }
        const float hfc20 = (float)Hi[2][0], hfc21 = (float)Hi[2][1], hfc22 = (float)Hi[2][2], hfc23 = (float)Hi[2][3];   /* :303-304 */
        double K[3][3] = {{view[0], 0, view[1]}, {0, view[0], view[2]}, {0, 0, 1}};      /* :450-455 */
        float pfc[3][4];
        for (int i = 0; i < 3; i++)                                              /* :307, Matrix::operator* matrix.cpp:396-418 */
            for (int j = 0; j < 4; j++) {
                double acc = 0.0;
                for (int k = 0; k < 3; k++) acc += K[i][k] * Hi[k][j];
                pfc[i][j] = (float)acc;
            }
        for (int32_t u = 0; u < width; u++)                                      /* :315-403 */
            for (int32_t v = 0; v < height; v++) {
                const int32_t addr = v * width + u;
                const float d = pD[addr];
                if (!(d > 0)) continue;
                const float x = pX[addr], y = pY[addr], z = pZ[addr];
                const float z2 = hfc20 * x + hfc21 * y + hfc22 * z + hfc23;
                int added = 0;
                if (((double)z2 > 0.1) && (z2 < max_dist)) {
                    const float w2 = pfc[2][0] * x + pfc[2][1] * y + pfc[2][2] * z + pfc[2][3];
                    const float qu = (pfc[0][0] * x + pfc[0][1] * y + pfc[0][2] * z + pfc[0][3]) / w2;
                    const float qv = (pfc[1][0] * x + pfc[1][1] * y + pfc[1][2] * z + pfc[1][3]) / w2;
                    /* (int32_t) of a float on x86-64 (cvttss2si): out of range and NaN give INT32_MIN */
                    const int32_t u2 = (qu >= -2147483648.0f && qu < 2147483648.0f) ? (int32_t)qu : INT32_MIN;
                    const int32_t v2 = (qv >= -2147483648.0f && qv < 2147483648.0f) ? (int32_t)qv : INT32_MIN;
                    if (u2 >= 0 && u2 < width && v2 >= 0 && v2 < height) {
                        const int32_t addr2 = v2 * width + u2;
                        const float d2 = cD[addr2];
                        if (d2 > 0) {
                            if ((double)(fabsf(x - cX[addr2]) + fabsf(y - cY[addr2]) + fabsf(z - cZ[addr2])) < 0.2) {   /* :359 */
                                cX[addr2] = (float)((cX[addr2] + x) / 2.0);
                                cY[addr2] = (float)((cY[addr2] + y) / 2.0);
                                cZ[addr2] = (float)((cZ[addr2] + z) / 2.0);
                                cI[addr2] = (float)((cI[addr2] + pI[addr]) / 2.0);
                                added = 1;
                            }
                        } else {
                            cX[addr2] = x; cY[addr2] = y; cZ[addr2] = z;
                            cI[addr2] = pI[addr];
                            cD[addr2] = 1;
                            added = 1;
                        }
                    }
                }
                if (added) pD[addr] = -1;                                        /* :383 */
                else {
                    float* o = points_prev + 4 * (size_t)(*n_prev)++;
                    o[0] = x; o[1] = y; o[2] = z; o[3] = pI[addr];
                }
            }
    }
    *n_curr = 0;
    for (int32_t u = 0; u < width; u++)                                          /* :414-430 */
        for (int32_t v = 0; v < height; v++) {
            const int32_t addr = v * width + u;
            if (cD[addr] > 0) {
                float* o = points_curr + 4 * (size_t)(*n_curr)++;
                o[0] = cX[addr]; o[1] = cY[addr]; o[2] = cZ[addr]; o[3] = cI[addr];
            }
        }
}

/* ------------------------------------------------------------------------------------------------
 * the feature filters of libviso2's Matcher, libviso2/src/filter.cpp:474-530, as
 * Matcher::computeFeatures calls them (matcher.cpp:799-801): sobel5x5 -> du, dv; blob5x5 -> f1; checkerboard5x5 -> f2.
 * The reference walks the image as ONE flat array of w*h elements (w = bytes per line, a multiple of 16): the
 * horizontal taps of a pixel near a row end come from the neighbouring row.  Plain scalar restatement:
 *   column pass (filter.cpp:291-349, :353-391): rows 2..h-3, zero elsewhere (memset)
 *     tv = i(-2) + 4 i(-1) + 6 i(0) + 4 i(+1) + i(+2)        th = i(-2) + 2 i(-1) - 2 i(+1) - i(+2)
 *     tc = i(-2) + i(-1) - i(+1) - i(+2)
 *   row pass, flat (filter.cpp:124-173, :180-222, :395-417): for j = 0, 1, ...
 *     du[j+2] = sat_u8(((tv[j] + 2 tv[j+1] - 2 tv[j+3] - tv[j+4]) >> 7) + 128)          j < w*h  (*)
 *     dv[j+2] = sat_u8(((th[j] + 4 th[j+1] + 6 th[j+2] + 4 th[j+3] + th[j+4]) >> 7) + 128)
 *     f2[j+2] = tc[j] + tc[j+1] - tc[j+3] - tc[j+4]                                      j < w*h - 8
 *   (*) the reference reads tv/th[w*h .. w*h+3] (beyond its buffers) for the last outputs and stores du/dv[w*h],
 *   [w*h+1]; here the temporaries read as 0 there and nothing is stored outside: du/dv[w*h-2 .. w*h) are not
 *   comparable with the reference.
 *   blob (filter.cpp:507-530): 2-d integral image in int32 (wrapping), then for p = 0 .. w*h-5-5w-1 (flat!)
 *     f1[p+3+3w] = (int16)(-(I[p+5+5w] - I[p+5] - I[p+5w] + I[p]) + 2 (I[p+4+4w] - I[p+4+w] - I[p+1+4w] + I[p+1+w])
 *                          + 7 in[p+3+3w])
 * Elements the reference leaves unwritten are 0. */
void oracle_matcher_filters(const uint8_t* in, int32_t w, int32_t h, uint8_t* du, uint8_t* dv, int16_t* f1, int16_t* f2)
{
    const size_t n = (size_t)w * h;
    int16_t* tv = (int16_t*)calloc(n + 8, sizeof(int16_t));
    int16_t* th = (int16_t*)calloc(n + 8, sizeof(int16_t));
    int16_t* tc = (int16_t*)calloc(n + 8, sizeof(int16_t));
    uint32_t* I = (uint32_t*)calloc(n, sizeof(uint32_t));
    memset(du, 0, n); memset(dv, 0, n); memset(f1, 0, 2 * n); memset(f2, 0, 2 * n);
    for (int32_t v = 2; v < h - 2; v++)
        for (int32_t u = 0; u < w; u++) {
            const size_t q = (size_t)v * w + u;
            const int a = in[q - 2 * (size_t)w], b = in[q - w], c = in[q], d = in[q + w], e = in[q + 2 * (size_t)w];
            tv[q] = (int16_t)(a + 4 * b + 6 * c + 4 * d + e);
            th[q] = (int16_t)(a + 2 * b - 2 * d - e);
            tc[q] = (int16_t)(a + b - d - e);
        }
    for (size_t j = 0; j + 2 < n; j++) {
        const int su = ((tv[j] + 2 * tv[j + 1] - 2 * tv[j + 3] - tv[j + 4]) >> 7) + 128;
        const int sv = ((th[j] + 4 * th[j + 1] + 6 * th[j + 2] + 4 * th[j + 3] + th[j + 4]) >> 7) + 128;
        du[j + 2] = (uint8_t)(su < 0 ? 0 : su > 255 ? 255 : su);
        dv[j + 2] = (uint8_t)(sv < 0 ? 0 : sv > 255 ? 255 : sv);
        if (j + 8 < n) f2[j + 2] = (int16_t)(tc[j] + tc[j + 1] - tc[j + 3] - tc[j + 4]);
    }
    for (int32_t v = 0; v < h; v++) {                       /* integral_image, filter.cpp:48-72 */
        uint32_t line = 0;
        for (int32_t u = 0; u < w; u++) {
            line += in[(size_t)v * w + u];
            I[(size_t)v * w + u] = (v ? I[(size_t)(v - 1) * w + u] : 0u) + line;
        }
    }
    if (n > 5 + 5 * (size_t)w)
        for (size_t p = 0; p + 5 + 5 * (size_t)w < n; p++) {
            const size_t W = (size_t)w;
            uint32_t r = 0u - (I[p + 5 + 5 * W] - I[p + 5] - I[p + 5 * W] + I[p]);
            r += 2u * (I[p + 4 + 4 * W] - I[p + 4 + W] - I[p + 1 + 4 * W] + I[p + 1 + W]);
            r += 7u * in[p + 3 + 3 * W];
            f1[p + 3 + 3 * W] = (int16_t)(uint16_t)r;
        }
    free(tv); free(th); free(tc); free(I);
}
