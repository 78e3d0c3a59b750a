/* oracle/orc_trace.c -- TEST INFRASTRUCTURE (CPU oracle).
 * RK4 propagator (restates Code/GeoAc/GeoAc.Solver.cpp:12-72) and the per-ray body of the `-prop` loops
 * (Code/GeoAc3D_main.cpp:226-304 and the four sibling mains), producing the records of include/geoac_b200.h.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "orc_eqsets.h"

static const orc_eqset* eqset_for(int variant) {
    switch (variant) {
        case GEOAC_2D: return &orc_eq_2d;
        case GEOAC_3D: return &orc_eq_3d;
        case GEOAC_GLOBAL: return &orc_eq_global;
        case GEOAC_3D_RNGDEP: return &orc_eq_3drngdep;
        case GEOAC_GLOBAL_RNGDEP: return &orc_eq_globalrngdep;
    }
    return 0;
}

/* GeoAc_Propagate_RK4, Solver.cpp:12-72.  sol is [rows][ORC_MAXEQ]; returns k+1 and *left_region. */
static int propagate_rk4(const orc_eqset* e, orc_ray* r, double (*sol)[ORC_MAXEQ], int step_limit, int* left_region) {
    const int n = r->eq_cnt;
    double t0[ORC_MAXEQ], t1[ORC_MAXEQ], t2[ORC_MAXEQ], t3[ORC_MAXEQ], t4[ORC_MAXEQ];
    double p1[ORC_MAXEQ], p2[ORC_MAXEQ], p3[ORC_MAXEQ];
    int k;
    *left_region = 0;
    for (k = 0; k < step_limit - 1; k++) {
        for (int i = 0; i < n; i++) t0[i] = sol[k][i];
        e->update(r, t0);
        double ds = e->set_ds(r, t0);
        for (int i = 0; i < n; i++) { t1[i] = ds * e->rhs(r, t0, i); p1[i] = sol[k][i] + t1[i] / 2.0; }
        e->update(r, p1);
        for (int i = 0; i < n; i++) { t2[i] = ds * e->rhs(r, p1, i); p2[i] = sol[k][i] + t2[i] / 2.0; }
        e->update(r, p2);
        for (int i = 0; i < n; i++) { t3[i] = ds * e->rhs(r, p2, i); p3[i] = sol[k][i] + t3[i]; }
        e->update(r, p3);
        for (int i = 0; i < n; i++) {
            t4[i] = ds * e->rhs(r, p3, i);
            sol[k + 1][i] = sol[k][i] + t1[i] / 6.0 + t2[i] / 3.0 + t3[i] / 3.0 + t4[i] / 6.0;
        }
        if (e->brk(r, sol[k + 1])) { *left_region = 1; break; }
        if (e->gnd(r, sol[k + 1])) { *left_region = 0; break; }
    }
    return k + 1;
}

/* orc_trace + the raypath rows of WriteRays=True (Code/GeoAc3D_main.cpp:249-262 and the sibling mains): with
 * path_stride > 0 one row { state[0..2], amplitude, attenuation, travel time, bounce, step } is appended per ray every
 * path_stride steps of the per-segment post pass (requires accum_per_segment semantics, which the mains use then); with
 * caus_cap > 0 also the WriteCaustics=True rows { state[0..2], travel time, bounce, step } where the Jacobian changes sign. */
int64_t orc_trace_paths(int variant, orc_atmo* atmo, const geoac_params* p, int64_t n_rays,
                        const double* theta, const double* phi, double* rec, int32_t* status, int32_t* n_steps,
                        int path_stride, int64_t path_cap, double* path, int32_t* path_rows,
                        int64_t caus_cap, double* caus, int32_t* caus_rows) {
    const orc_eqset* e = eqset_for(variant);
    if (!e || !atmo || !p) return -1;
    const int n_rec = p->bounces + 1;
    const int64_t n_slots = n_rays * n_rec;
    const int step_limit = (int)(p->ray_limit * (int)(1.0 / (p->ds_min * 10)));      /* Solver.cpp:14 */
    double (*sol)[ORC_MAXEQ] = (double (*)[ORC_MAXEQ])calloc((size_t)step_limit + 2, sizeof(double[ORC_MAXEQ]));
    if (!sol) return -1;
    for (int64_t i = 0; i < n_slots * GEOAC_NFIELDS; i++) rec[i] = 0.0;
    for (int64_t i = 0; i < n_slots; i++) { status[i] = GEOAC_ST_NONE; n_steps[i] = 0; }
    atmo->z_grnd = p->z_grnd; atmo->tweak_abs = p->tweak_abs;
    const int per_bounce_zmax = (variant == GEOAC_3D_RNGDEP || variant == GEOAC_GLOBAL_RNGDEP);  /* App. A-3 */
    const int seg_mode = (variant == GEOAC_2D) ? 1 : p->accum_per_segment;                      /* App. A-2 */
    if ((path_stride > 0 || caus_cap > 0) && !seg_mode) return -1;

    orc_ray ray; memset(&ray, 0, sizeof ray);
    ray.atmo = atmo; ray.prm = p; ray.calc_amp = p->calc_amp;
    ray.eq_cnt = p->calc_amp ? e->eq_amp : e->eq_noamp;
    int64_t total = 0;
    for (int64_t iray = 0; iray < n_rays; iray++) {
        ray.theta = theta[iray]; ray.phi = phi[iray];
        e->init(&ray, sol[0]);
        double tt = 0.0, att = 0.0, zmax = 0.0;
        int32_t prow = 0, crow = 0, crow_b = 0;
        for (int b = 0; b < n_rec; b++) {
            int64_t slot = iray * n_rec + b;
            int left; int k = propagate_rk4(e, &ray, sol, step_limit, &left);
            total += k;
            if (seg_mode) {
                double D = 0.0, D_prev = 0.0;
                crow_b = 0;
                if (caus_cap > 0) D_prev = e->jacobian(&ray, sol[1]);              /* Code/GeoAc3D_main.cpp:245 */
                for (int m = 1; m < k; m++) {
                    if (caus_cap > 0) D = e->jacobian(&ray, sol[m]);
                    e->tt_seg(&ray, sol[m - 1], sol[m], &tt); e->sb_seg(&ray, sol[m - 1], sol[m], &att);
                    if (path_stride > 0 && m % path_stride == 0) {
                        if (prow < path_cap) {
                            double* row = path + (iray * path_cap + prow) * GEOAC_PATH_NF;
                            row[0] = sol[m][0]; row[1] = sol[m][1]; row[2] = sol[m][2];
                            row[3] = p->calc_amp ? e->amplitude(&ray, sol[m]) : 0.0;
                            row[4] = att; row[5] = tt; row[6] = (double)b; row[7] = (double)m;
                        }
                        prow++;
                    }
                    if (caus_cap > 0 && D * D_prev < 0.0) {                        /* :263-268 */
                        if (crow < caus_cap) {
                            double* row = caus + (iray * caus_cap + crow) * GEOAC_CAUSTIC_NF;
                            row[0] = sol[m][0]; row[1] = sol[m][1]; row[2] = sol[m][2]; row[3] = tt; row[4] = (double)b; row[5] = (double)m;
                        }
                        crow++; crow_b++;
                    }
                    if (caus_cap > 0) D_prev = D;
                }
            } else {
                double t = 0.0, a = 0.0;
                for (int m = 0; m < k; m++) e->tt_seg(&ray, sol[m], sol[m + 1], &t);
                for (int m = 0; m < k; m++) e->sb_seg(&ray, sol[m], sol[m + 1], &a);
                tt += t; att += a;
            }
            n_steps[slot] = k;
            if (path_stride > 0) path_rows[iray] = prow;
            if (caus_cap > 0) caus_rows[iray] = crow;
            if (left) { status[slot] = GEOAC_ST_BREAK; break; }
            if (k >= step_limit) { status[slot] = GEOAC_ST_LIMIT; break; }
            status[slot] = GEOAC_ST_ARRIVAL;
            if (per_bounce_zmax) zmax = 0.0;
            for (int m = 0; m < k; m++) zmax = fmax(zmax, e->altitude(&ray, sol[m]));
            for (int i = 0; i < ray.eq_cnt; i++) rec[(int64_t)i * n_slots + slot] = sol[k][i];
            rec[(int64_t)GEOAC_F_TRAVELTIME * n_slots + slot] = tt;
            rec[(int64_t)GEOAC_F_ATTEN * n_slots + slot] = att;
            rec[(int64_t)GEOAC_F_TURNHEIGHT * n_slots + slot] = zmax;
            rec[(int64_t)GEOAC_F_AMPLITUDE * n_slots + slot] = p->calc_amp ? e->amplitude(&ray, sol[k]) : 0.0;
            rec[(int64_t)GEOAC_F_JACOBIAN * n_slots + slot] = p->calc_amp ? e->jacobian(&ray, sol[k]) : 0.0;
            rec[(int64_t)GEOAC_F_CAUSTICS * n_slots + slot] = (caus_cap > 0) ? (double)crow_b : -1.0;
            double incl, baz, aux, margin;
            e->finish(&ray, sol[k - 1], sol[k], tt, &incl, &baz, &aux, &margin);
            rec[(int64_t)GEOAC_F_INCLINATION * n_slots + slot] = incl;
            rec[(int64_t)GEOAC_F_BACKAZ * n_slots + slot] = baz;
            rec[(int64_t)GEOAC_F_AUX * n_slots + slot] = aux;
            rec[(int64_t)GEOAC_F_MARGIN * n_slots + slot] = margin;
            if (b + 1 < n_rec) {
                double y0[ORC_MAXEQ];
                e->reflect(&ray, sol[k - 2], sol[k - 1], sol[k], y0);
                for (int i = 0; i < ray.eq_cnt; i++) sol[0][i] = y0[i];
            }
        }
    }
    free(sol);
    return total;
}

int64_t orc_trace(int variant, orc_atmo* atmo, const geoac_params* p, int64_t n_rays,
                  const double* theta, const double* phi, double* rec, int32_t* status, int32_t* n_steps) {
    return orc_trace_paths(variant, atmo, p, n_rays, theta, phi, rec, status, n_steps, 0, 0, 0, 0, 0, 0, 0);
}

/* GeoAc_SetPropRegion: G2S_Spline1D.cpp:22-28, G2S_GlobalSpline1D.cpp:22-30; grids: orc_mspline.c */
void orc_mspline_region(const orc_atmo* a, geoac_params* p);
void orc_set_prop_region(int variant, const orc_atmo* a, geoac_params* p) {
    if (a->kind == 0) { p->vert_limit = a->vmax; p->range_limit = 10000.0; }
    else if (a->kind == 1) {
        p->vert_limit = a->vmax; p->range_limit = 1500.0;
        p->box_min[0] = -ORC_PI / 2.0; p->box_max[0] = ORC_PI / 2.0; p->box_min[1] = -ORC_PI; p->box_max[1] = ORC_PI;
    } else orc_mspline_region(a, p);
    (void)variant;
}

int geoac_default_params_oracle(int variant, geoac_params* p) {
    memset(p, 0, sizeof *p);
    p->ds_min = 0.001; p->ds_max = 0.5;
    p->ray_limit = (variant == GEOAC_GLOBAL || variant == GEOAC_GLOBAL_RNGDEP) ? 10000.0 : 5000.0;
    p->vert_limit = 200.0; p->range_limit = 2000.0;
    p->z_grnd = 0.0; p->tweak_abs = 0.3; p->freq = 0.1;
    p->bounces = 2; p->calc_amp = 1; p->accum_per_segment = (variant == GEOAC_2D);
    if (variant == GEOAC_GLOBAL) { p->src[1] = 30.0 * ORC_PI / 180.0; p->src[2] = 0.0; }
    return 0;
}
