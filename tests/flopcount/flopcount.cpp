// tests/flopcount/flopcount.cpp -- TEST INFRASTRUCTURE (never part of libgeoac_b200.so): compiles the per-ray device code
// with the op-counting scalar of counted.hpp in place of `double` and traces a few rays per variant, printing the
// ALGORITHMIC FP64 operations per RK4 step of the de-duplicated formulation: each libm call counts 1
// (GEOAC_COUNT_FLOPS makes g_exp_n / g_rcp / g_rsqrt call the counted exp / divide / sqrt instead of their polynomial
// and Newton expansions, which are an implementation detail of the device build).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../include/geoac_b200.h"
#include "counted.hpp"
using cnt::Cnt;
using cnt::fma; using cnt::sqrt; using cnt::exp; using cnt::sin; using cnt::cos; using cnt::tan; using cnt::asin; using cnt::cbrt;
using cnt::atan2; using cnt::pow; using cnt::sincos; using cnt::fabs; using cnt::floor; using cnt::fmax; using cnt::fmin; using cnt::ldexp;
using cnt::log10;
namespace std { inline cnt::Cnt pow(cnt::Cnt a, int b) { return cnt::pow(a, b); } }      // host_tables.hpp spells std::pow
#define GEOAC_COUNT_FLOPS 1
#define double Cnt
#include "../../geoac_b200/csrc/core.cuh"
#include "../../geoac_b200/csrc/eq_cartesian.cuh"
#include "../../geoac_b200/csrc/eq_global.cuh"
#include "../../geoac_b200/csrc/eq_rngdep.cuh"
#include "../../geoac_b200/csrc/trace_kernel.cuh"
#include "../../geoac_b200/csrc/host_tables.hpp"
#undef double

using namespace geoac;

static void base_consts(LaunchConsts& L, int variant, const geoac_params* p) {
    std::memset((void*)&L, 0, sizeof L);
    L.ds_min = p->ds_min; L.ds_max = p->ds_max; L.vert_limit = p->vert_limit; L.range_limit = p->range_limit;
    L.z_grnd = p->z_grnd; L.tweak_abs = p->tweak_abs; L.freq = p->freq;
    for (int i = 0; i < 2; i++) { L.box_min[i] = p->box_min[i]; L.box_max[i] = p->box_max[i]; }
    for (int i = 0; i < 3; i++) L.src[i] = p->src[i];
    L.bounces = p->bounces; L.calc_amp = p->calc_amp;
    L.seg_mode = (variant == GEOAC_2D) ? 1 : (p->accum_per_segment ? 1 : 0);
    L.step_limit = (int)(p->ray_limit * (int)(1.0 / (p->ds_min * 10)));
    L.per_bounce_zmax = (variant == GEOAC_3D_RNGDEP || variant == GEOAC_GLOBAL_RNGDEP);
}

template <class EQ>
static void run(const LaunchConsts& L, const typename EQ::Atmo& T, long n, const double* th, const double* ph, int n_rec, const char* name) {
    std::vector<Cnt> rec((size_t)GEOAC_NFIELDS * n * n_rec), prev(EQ::NEQ), work(2 * EQ::NEQ);
    std::vector<int32_t> status(n * n_rec), nsteps(n * n_rec);
    RecOut o; o.rec = rec.data(); o.status = status.data(); o.n_steps = nsteps.data(); o.n_rec = n_rec; o.n_slots = n * n_rec;
    o.path = nullptr; o.path_rows = nullptr; o.path_stride = 0; o.path_cap = 0; o.caus = nullptr; o.caus_rows = nullptr; o.caus_cap = 0;
    cnt::tally() = cnt::Tally();
    long steps = 0;
    for (long i = 0; i < n; i++) {
        LaneD<EQ> ld; LaneI<EQ> li;
        lane_start<EQ>(ld, li, L, T, i, Cnt(th[i]), Cnt(ph[i]));
        while (lane_advance<EQ>(ld, li, L, T, prev.data(), 1, o, work.data())) steps++;
        steps++;
    }
    const cnt::Tally& t = cnt::tally();
    std::printf("{\"variant\": \"%s\", \"rays\": %ld, \"rk4_steps\": %ld, \"flops_per_step\": %.1f, \"add\": %.1f, \"mul\": %.1f, \"fma\": %.1f, \"div\": %.1f, "
                "\"sqrt\": %.1f, \"transcendental\": %.1f}\n", name, n, steps, (double)cnt::flops() / steps, (double)t.add / steps, (double)t.mul / steps,
                (double)t.fma / steps, (double)t.div / steps, (double)t.sqrt / steps, (double)t.trans / steps);
}

// 1-D variants: table = n records of TAB_NARR doubles (as geoac_set_atmosphere_1d builds it)
extern "C" int flopcount_1d(int variant, const geoac_params* p, int n, const double* table, long n_rays, const double* th, const double* ph) {
    std::vector<Cnt> tab(table, table + (size_t)n * TAB_NARR);
    Table1D T; T.base = tab.data(); T.n = n; T.xmin = tab[TAB_X]; T.xmax = tab[(size_t)(n - 1) * TAB_NARR + TAB_X]; T.jump_scale = 0.0;
    LaunchConsts L; base_consts(L, variant, p);
    if (variant == GEOAC_2D || variant == GEOAC_3D) L.src[2] = std::max(p->z_grnd, p->src[2]); else L.src[0] = std::max(p->z_grnd, p->src[0]);
    fill_launch_consts_1d(L, T, variant);
    // absorption through the per-interval polynomials, as the product does (core.cuh); building them is set-up, not counted (run() resets the tally)
    std::vector<Cnt> sbp((size_t)(n - 1) * SBP_STRIDE);
    for (int k = 0; k < n - 1; k++) sbpoly_build_interval(L, T, variant == GEOAC_GLOBAL, k, &sbp[(size_t)k * SBP_STRIDE]);
    T.sbpoly = sbp.data();
    const int n_rec = p->bounces + 1;
    switch (variant) {
        case GEOAC_2D: run<Eq2D<true>>(L, T, n_rays, th, ph, n_rec, "2d"); return 0;
        case GEOAC_3D: run<Eq3D<true>>(L, T, n_rays, th, ph, n_rec, "3d"); return 0;
        case GEOAC_GLOBAL: run<EqGlobal<true>>(L, T, n_rays, th, ph, n_rec, "global"); return 0;
    }
    return -1;
}

extern "C" int flopcount_3d(int variant, const geoac_params* p, int n0, int n1, int nz, const double* ax0, const double* ax1, const double* axz,
                            const double* Tf, const double* uf, const double* vf, const double* rhof, long n_rays, const double* th, const double* ph) {
    const bool glob = variant == GEOAC_GLOBAL_RNGDEP;
    const size_t nodes = (size_t)n0 * n1 * nz;
    auto lift = [](const double* a, size_t n) { return std::vector<Cnt>(a, a + n); };
    std::vector<Cnt> a0 = lift(ax0, n0), a1 = lift(ax1, n1), az = lift(axz, nz), T_ = lift(Tf, nodes), u_ = lift(uf, nodes), v_ = lift(vf, nodes), r_ = lift(rhof, nodes);
    std::vector<Cnt> z, tuv, rh;
    build_grid_tables(glob, n0, n1, nz, a0.data(), a1.data(), az.data(), T_.data(), u_.data(), v_.data(), r_.data(), z, tuv, rh);
    std::vector<Cnt> r0, r1, rz;
    build_axis_records(a0.data(), n0, r0); build_axis_records(a1.data(), n1, r1); build_axis_records(z.data(), nz, rz);
    Grid3D g; g.tuv = tuv.data(); g.rho = rh.data(); g.ax0 = r0.data(); g.ax1 = r1.data(); g.axz = rz.data(); g.n0 = n0; g.n1 = n1; g.nz = nz;
    g.amin = a0[0]; g.amax = a0[n0 - 1]; g.bmin = a1[0]; g.bmax = a1[n1 - 1]; g.zmin = z[0]; g.zmax = z[nz - 1];
    static Cnt scratch[MS_SCRATCH]; g.scratch = scratch; g.role = 0; g.nrole = 1; g.glane0 = 0; g.gmask = 0;
    LaunchConsts L; base_consts(L, variant, p);
    if (!glob) L.src[2] = std::max(p->z_grnd, p->src[2]); else L.src[0] = std::max(p->z_grnd, p->src[0]);
    fill_launch_consts_3d(L, g, variant);
    const int n_rec = p->bounces + 1;
    if (!glob) run<Eq3DRD<true>>(L, g, n_rays, th, ph, n_rec, "3drngdep"); else run<EqGlobalRD<true>>(L, g, n_rays, th, ph, n_rec, "globalrngdep");
    return 0;
}
