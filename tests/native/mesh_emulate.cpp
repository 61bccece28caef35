// TEST INFRASTRUCTURE.  Runs the phases of stereo-vision_b200/csrc/mesh_core.h (the code the k_mesh kernels
// execute between CTA barriers) on the CPU: every phase is executed for tid = 0..nthr-1 in a scrambled order,
// barriers are the boundaries between the loops.  tests/test_mesh_core.py compares the result with the
// sequential host stage (elas_b200_host_stage), which is pinned to the reference by tests/test_oracle.py.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../stereo-vision_b200/csrc/mesh_core.h"

using namespace elasb::mesh;

namespace {
struct Emu {
    int nthr; std::vector<int> order;
    Emu(int n, unsigned seed) : nthr(n), order(n) {
        std::iota(order.begin(), order.end(), 0);
        for (int i = n - 1; i > 0; i--) { seed = seed * 1664525u + 1013904223u; std::swap(order[i], order[(seed >> 8) % (i + 1)]); }
    }
    template <class F> void phase(F f) { for (int t : order) f(t, nthr); }
};
void inclusive_scan(const int32_t* in, int32_t* out, int n) { int s = 0; for (int i = 0; i < n; i++) { s += in[i]; out[i] = s; } }
}  // namespace

extern "C" {

// dcan [Hc][Wc] in/out; dcan_incon out; support (n x 3) out; returns n_support
int mesh_emulate_lattice(int Wc, int Hc, int step, int win, int thr, int need, int16_t* dcan, int16_t* dcan_incon,
                         int32_t* support, int nthr, unsigned seed, int* rounds_out)
{
    Emu emu(nthr, seed);
    std::vector<int16_t> pad((size_t)lat_elems(Wc, Hc));
    Lattice L{Wc, Hc, Wc + 2 * kPadC, step, pad.data()};
    emu.phase([&](int t, int n) { lattice_load(L, dcan, t, n); });
    int rounds = 0;
    auto or16 = [](int16_t* p, int bits) { *p = (int16_t)(*p | bits); };
    auto atomic_add = [](int* p, int v) { const int old = *p; *p += v; return old; };
    std::vector<int32_t> sup_cnt((size_t)Wc * Hc), lists[2] = {std::vector<int32_t>((size_t)Wc * Hc), std::vector<int32_t>((size_t)Wc * Hc)};
    emu.phase([&](int t, int n) { for (int i = t; i < Wc * Hc; i += n) sup_cnt[i] = incon_count0(dcan, Wc, Hc, i % Wc, i / Wc, win, thr); });
    int n_cur = 0;
    emu.phase([&](int t, int n) { incon_seed(L, sup_cnt.data(), need, lists[0].data(), &n_cur, t, n, atomic_add); });
    for (int cur = 0; n_cur > 0; cur ^= 1) {
        int n_next = 0;
        emu.phase([&](int t, int n) { incon_propagate(L, sup_cnt.data(), win, thr, need, lists[cur].data(), n_cur, lists[cur ^ 1].data(), &n_next, t, n, atomic_add, or16); });
        n_cur = n_next;
        rounds++;
    }
    if (rounds_out) *rounds_out = rounds;
    emu.phase([&](int t, int n) { incon_finish(L, dcan_incon, t, n); });
    emu.phase([&](int t, int n) { redundant_pass(L, true, t, n); });
    emu.phase([&](int t, int n) { redundant_pass(L, false, t, n); });
    std::vector<int32_t> cnt(Wc), off(Wc);
    emu.phase([&](int t, int n) { support_count(L, cnt.data(), t, n); });
    inclusive_scan(cnt.data(), off.data(), Wc);
    const int total = off[Wc - 1];
    for (int i = 0; i < Wc; i++) off[i] -= cnt[i];
    emu.phase([&](int t, int n) { support_write(L, off.data(), support, t, n); });
    emu.phase([&](int t, int n) { lattice_store(L, dcan, t, n); });
    return total;
}

// support (n x 3); tri_out (cap x 3); units_out (2 ints each); returns number of triangles, *n_units
int mesh_emulate_delaunay(const int32_t* support, int n, int right_image, int W, int H, int band_rows, int unit_cap,
                          int32_t* tri_out, int32_t* units_out, int* n_units, int32_t* overflow_out, int* n_overflow,
                          int nthr, unsigned seed)
{
    Emu emu(nthr, seed);
    std::vector<int32_t> x(n), y(n);
    for (int i = 0; i < n; i++) { x[i] = right_image ? support[3 * i] - support[3 * i + 2] : support[3 * i]; y[i] = support[3 * i + 1]; }
    // the two sorted id lists (a CTA-wide bitonic sort of (key, id) on the device; keys are unique)
    std::vector<int32_t> xs(n), ys(n), posx(n), posy(n), seg_lo(n), seg_n(n), side(n), scan(n), tmp(n);
    std::iota(xs.begin(), xs.end(), 0); std::iota(ys.begin(), ys.end(), 0);
    std::sort(xs.begin(), xs.end(), [&](int a, int b) { return x[a] != x[b] ? x[a] < x[b] : y[a] < y[b]; });
    std::sort(ys.begin(), ys.end(), [&](int a, int b) { return y[a] != y[b] ? y[a] < y[b] : x[a] < x[b]; });
    for (int i = 1; i < n; i++) if (x[xs[i]] == x[xs[i - 1]] && y[xs[i]] == y[xs[i - 1]]) return -1;   // duplicates: not this path
    Order o{n, xs.data(), ys.data(), posx.data(), posy.data(), seg_lo.data(), seg_n.data(), side.data(), scan.data(), tmp.data()};
    emu.phase([&](int t, int k) { order_init(o, t, k); });
    for (int axis = 0;; axis ^= 1) {
        bool cutting = false;
        emu.phase([&](int t, int k) { cutting |= order_flags(o, axis, t, k); });
        if (!cutting) break;
        inclusive_scan(side.data(), scan.data(), n);
        emu.phase([&](int t, int k) { order_scatter(o, axis, t, k); });
        emu.phase([&](int t, int k) { order_commit(o, axis, t, k); });
    }
    std::vector<int32_t> nbr(4 * (size_t)(2 * n)), vtx(4 * (size_t)(2 * n)), hull(2 * (size_t)n);
    std::vector<uint32_t> xy(n);
    for (int i = 0; i < n; i++) xy[i] = ((uint32_t)x[i] << 16) | (uint32_t)y[i];
    Mesh m{n, xy.data(), xs.data(), nbr.data(), vtx.data(), hull.data()};
    for (int depth = tree_depth(n); depth >= 0; depth--)
        emu.phase([&](int t, int k) { triangulate_depth(m, depth, t, k); });
    const int nt_pool = 2 * n - 2;
    std::vector<int32_t> flag(nt_pool), fscan(nt_pool);
    emu.phase([&](int t, int k) { real_flags(m, flag.data(), t, k); });
    inclusive_scan(flag.data(), fscan.data(), nt_pool);
    const int nt = fscan[nt_pool - 1];
    emu.phase([&](int t, int k) { write_triangles(m, flag.data(), fscan.data(), tri_out, t, k); });
    std::vector<int32_t> cnt(nt), cscan(nt);
    emu.phase([&](int t, int k) { unit_counts(x.data(), y.data(), tri_out, nt, W, H, band_rows, cnt.data(), t, k); });
    inclusive_scan(cnt.data(), cscan.data(), nt);
    emu.phase([&](int t, int k) { write_units(x.data(), y.data(), tri_out, nt, W, H, band_rows, right_image, cnt.data(), cscan.data(), unit_cap, units_out, t, k); });
    int nu = 0, no = 0;
    for (int t = 0; t < nt; t++) {
        if (cnt[t] == 0) continue;
        if (unit_overflows(cscan.data(), t, unit_cap)) overflow_out[no++] = t; else nu = cscan[t];
    }
    *n_units = nu; *n_overflow = no;
    return nt;
}

}  // extern "C"
