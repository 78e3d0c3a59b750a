// tests/host_emul/host_emul.cpp -- TEST-ONLY debugging aid (never part of libgeoac_b200.so, never a fallback).
// Compiles the per-ray state machine of geoac_b200/csrc (lane_advance<EQ> and the equation sets, all GEOAC_HD) with
// g++ so that the de-duplicated device math can be checked against the oracle in the build container, where there
// is no GPU.  The real parity gate is tests/test_gpu_parity.py on a B200.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include "../../geoac_b200/csrc/core.cuh"
#include "../../geoac_b200/csrc/eq_cartesian.cuh"
#if __has_include("../../geoac_b200/csrc/eq_global.cuh")
#include "../../geoac_b200/csrc/eq_global.cuh"
#define HAVE_GLOBAL 1
#endif
#include "../../geoac_b200/csrc/eq_rngdep.cuh"
#include "../../geoac_b200/csrc/trace_kernel.cuh"
#include "../../geoac_b200/csrc/host_tables.hpp"

using namespace geoac;

template <class EQ, bool PATHS>
static long run_mode(const LaunchConsts& L, const typename EQ::Atmo& T, long n, const double* th, const double* ph, RecOut o) {
    long steps = 0;
    std::vector<double> prev(EQ::NEQ, 0.0), work(2 * EQ::NEQ, 0.0);
    for (long i = 0; i < n; i++) {
        LaneD<EQ> ld; LaneI<EQ> li;
        lane_start<EQ>(ld, li, L, T, i, th[i], ph[i]);
        while (lane_advance<EQ, PATHS>(ld, li, L, T, prev.data(), 1, o, work.data())) steps++;
        steps++;
    }
    return steps;
}
template <class EQ>
static long run(const LaunchConsts& L, const typename EQ::Atmo& T, long n, const double* th, const double* ph, RecOut o) {
    return (o.path_stride > 0 || o.caus_cap > 0) ? run_mode<EQ, true>(L, T, n, th, ph, o) : run_mode<EQ, false>(L, T, n, th, ph, o);
}

extern "C" long emul_trace_1d(int variant, const geoac_params* p, int n, const double* table, long n_rays,
                              const double* th, const double* ph, double* rec, int32_t* status, int32_t* n_steps,
                              int path_stride, long path_cap, double* path, int32_t* path_rows, long caus_cap, double* caus, int32_t* caus_rows) {
    Table1D T; T.base = table; T.n = n; T.xmin = table[TAB_X]; T.xmax = table[(size_t)(n - 1) * TAB_NARR + TAB_X]; T.jump_scale = 0.0;
    LaunchConsts L; std::memset(&L, 0, sizeof L);
    L.ds_min = p->ds_min; L.ds_max = p->ds_max; L.vert_limit = p->vert_limit; L.range_limit = p->range_limit;
    L.z_grnd = p->z_grnd; L.tweak_abs = p->tweak_abs; L.freq = p->freq;
    for (int i = 0; i < 2; i++) { L.box_min[i] = p->box_min[i]; L.box_max[i] = p->box_max[i]; }
    for (int i = 0; i < 3; i++) L.src[i] = p->src[i];
    if (variant == GEOAC_2D || variant == GEOAC_3D) L.src[2] = std::max(p->z_grnd, p->src[2]); else L.src[0] = std::max(p->z_grnd, p->src[0]);
    L.bounces = p->bounces; L.calc_amp = p->calc_amp;
    L.seg_mode = (variant == GEOAC_2D) ? 1 : (p->accum_per_segment ? 1 : 0);
    L.step_limit = (int)(p->ray_limit * (int)(1.0 / (p->ds_min * 10)));
    L.per_bounce_zmax = 0;
    fill_launch_consts_1d(L, T, variant);
    std::vector<double> sbp;                                  // GEOAC_EMUL_SBPOLY=1: absorption through the per-interval polynomials, as on the device
    if (const char* e = std::getenv("GEOAC_EMUL_SBPOLY")) if (std::atoi(e) != 0) {
        sbp.resize((size_t)(n - 1) * SBP_STRIDE);
        long flagged = 0;
        for (int k = 0; k < n - 1; k++) { sbpoly_build_interval(L, T, variant == GEOAC_GLOBAL, k, &sbp[(size_t)k * SBP_STRIDE]); flagged += sbp[(size_t)k * SBP_STRIDE + 7] != 0.0; }
        if (std::atoi(e) > 1) std::fprintf(stderr, "sbpoly: %ld of %d intervals flagged for exact evaluation\n", flagged, n - 1);
        T.sbpoly = sbp.data();
    }
    RecOut o; o.rec = rec; o.status = status; o.n_steps = n_steps; o.n_rec = p->bounces + 1; o.n_slots = n_rays * o.n_rec;
    o.path = path; o.path_rows = path_rows; o.path_stride = path_stride; o.path_cap = path_cap;
    o.caus = caus; o.caus_rows = caus_rows; o.caus_cap = caus_cap;
    std::fill(rec, rec + (size_t)GEOAC_NFIELDS * o.n_slots, 0.0);
    std::fill(status, status + o.n_slots, 0); std::fill(n_steps, n_steps + o.n_slots, 0);
    const bool amp = p->calc_amp != 0;
    switch (variant) {
        case GEOAC_2D: return amp ? run<Eq2D<true>>(L, T, n_rays, th, ph, o) : run<Eq2D<false>>(L, T, n_rays, th, ph, o);
        case GEOAC_3D: return amp ? run<Eq3D<true>>(L, T, n_rays, th, ph, o) : run<Eq3D<false>>(L, T, n_rays, th, ph, o);
#ifdef HAVE_GLOBAL
        case GEOAC_GLOBAL: return amp ? run<EqGlobal<true>>(L, T, n_rays, th, ph, o) : run<EqGlobal<false>>(L, T, n_rays, th, ph, o);
#endif
    }
    return -1;
}

// range-dependent variants: fields dense [n0][n1][nz] exactly as geoac_set_atmosphere_3d receives them
extern "C" long emul_trace_3d(int variant, const geoac_params* p, int n0, int n1, int nz, const double* ax0, const double* ax1, const double* axz,
                              const double* Tf, const double* uf, const double* vf, const double* rhof, long n_rays,
                              const double* th, const double* ph, double* rec, int32_t* status, int32_t* n_steps,
                              int path_stride, long path_cap, double* path, int32_t* path_rows, long caus_cap, double* caus, int32_t* caus_rows) {
    const bool glob = variant == GEOAC_GLOBAL_RNGDEP;
    std::vector<double> z, tuv, rh;
    build_grid_tables(glob, n0, n1, nz, ax0, ax1, axz, Tf, uf, vf, rhof, z, tuv, rh);
    std::vector<double> r0, r1, rz;
    build_axis_records(ax0, n0, r0); build_axis_records(ax1, n1, r1); build_axis_records(z.data(), nz, rz);
    Grid3D g; g.tuv = tuv.data(); g.rho = rh.data(); g.ax0 = r0.data(); g.ax1 = r1.data(); g.axz = rz.data(); g.n0 = n0; g.n1 = n1; g.nz = nz;
    g.amin = ax0[0]; g.amax = ax0[n0 - 1]; g.bmin = ax1[0]; g.bmax = ax1[n1 - 1]; g.zmin = z[0]; g.zmax = z[nz - 1];
    static thread_local double scratch[MS_SCRATCH]; g.scratch = scratch; g.role = 0; g.nrole = 1; g.glane0 = 0; g.gmask = 0;
    LaunchConsts L; std::memset(&L, 0, sizeof L);
    L.ds_min = p->ds_min; L.ds_max = p->ds_max; L.vert_limit = p->vert_limit; L.range_limit = p->range_limit;
    L.z_grnd = p->z_grnd; L.tweak_abs = p->tweak_abs; L.freq = p->freq;
    for (int i = 0; i < 2; i++) { L.box_min[i] = p->box_min[i]; L.box_max[i] = p->box_max[i]; }
    for (int i = 0; i < 3; i++) L.src[i] = p->src[i];
    if (!glob) L.src[2] = std::max(p->z_grnd, p->src[2]); else L.src[0] = std::max(p->z_grnd, p->src[0]);
    L.bounces = p->bounces; L.calc_amp = p->calc_amp;
    L.seg_mode = p->accum_per_segment ? 1 : 0;
    L.step_limit = (int)(p->ray_limit * (int)(1.0 / (p->ds_min * 10)));
    L.per_bounce_zmax = 1;
    fill_launch_consts_3d(L, g, variant);
    RecOut o; o.rec = rec; o.status = status; o.n_steps = n_steps; o.n_rec = p->bounces + 1; o.n_slots = n_rays * o.n_rec;
    o.path = path; o.path_rows = path_rows; o.path_stride = path_stride; o.path_cap = path_cap;
    o.caus = caus; o.caus_rows = caus_rows; o.caus_cap = caus_cap;
    std::fill(rec, rec + (size_t)GEOAC_NFIELDS * o.n_slots, 0.0);
    std::fill(status, status + o.n_slots, 0); std::fill(n_steps, n_steps + o.n_slots, 0);
    const bool amp = p->calc_amp != 0;
    if (!glob) return amp ? run<Eq3DRD<true>>(L, g, n_rays, th, ph, o) : run<Eq3DRD<false>>(L, g, n_rays, th, ph, o);
    return amp ? run<EqGlobalRD<true>>(L, g, n_rays, th, ph, o) : run<EqGlobalRD<false>>(L, g, n_rays, th, ph, o);
}
