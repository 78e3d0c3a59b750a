/* oracle/geoac_oracle.h -- TEST INFRASTRUCTURE (CPU oracle), never linked into or called by the product.
 *
 * Plain-C restatement of the reference's hot path (GeoAc_Propagate_RK4 + equation sets + atmosphere splines +
 * Sutherland-Bass absorption + the per-ray body of the `-prop` loops).  It keeps the reference's floating-point
 * expression trees so that, compiled with `-O2 -ffp-contract=off`, it is BIT-IDENTICAL to the unmodified reference
 * built by oracle/Makefile (`make ref`) -- tests/test_oracle_vs_ref.py pins that on every golden vector.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 */
#ifndef GEOAC_ORACLE_H_
#define GEOAC_ORACLE_H_

#include <stdint.h>
#include "../include/geoac_b200.h"   /* shares only the POD types/enums of the public ABI */

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAXEQ 18

/* natural cubic spline in "slopes" form, Code/Atmo/G2S_Spline1D.h:43-49 */
typedef struct orc_spline1d {
    int n;
    int accel;              /* cursor of the previous look-up (persists across rays, like the reference) */
    const double* x;
    const double* f;
    double* slopes;
} orc_spline1d;

void   orc_spline1d_set_slopes(orc_spline1d* s);
double orc_spline1d_f(double x, orc_spline1d* s);
double orc_spline1d_df(double x, orc_spline1d* s);
double orc_spline1d_ddf(double x, orc_spline1d* s);

/* atmosphere behind the Atmo_State.h API (Code/Atmo/Atmo_State.h:11-36) */
typedef struct orc_atmo orc_atmo;
struct orc_atmo {
    int kind;               /* 0: 1-D Cartesian, 1: 1-D Global, 2: 3-D Cartesian grid, 3: 3-D Global grid */
    int vert_index;         /* derivative index of the vertical coordinate: 2 Cartesian, 0 Global */
    double vmin, vmax;      /* clamp range of the vertical coordinate */
    double r_earth;
    double z_grnd;
    double tweak_abs;
    orc_spline1d T, U, V, RHO;
    double *xv, *Tv, *Uv, *Vv, *RHOv;       /* owned copies for 1-D */
    void* grid;                              /* range-dependent tables (orc_mspline.c) */
    double (*c)(orc_atmo*, double, double, double);
    double (*c_diff)(orc_atmo*, double, double, double, int);
    double (*c_ddiff)(orc_atmo*, double, double, double, int, int);
    double (*u)(orc_atmo*, double, double, double);
    double (*u_diff)(orc_atmo*, double, double, double, int);
    double (*u_ddiff)(orc_atmo*, double, double, double, int, int);
    double (*v)(orc_atmo*, double, double, double);
    double (*v_diff)(orc_atmo*, double, double, double, int);
    double (*v_ddiff)(orc_atmo*, double, double, double, int, int);
    double (*rho)(orc_atmo*, double, double, double);
};

orc_atmo* orc_atmo1d_create(int global, int n, const double* z, const double* T, const double* u,
                            const double* v, const double* rho);
orc_atmo* orc_atmo3d_create(int global, int n0, int n1, int nz, const double* ax0, const double* ax1,
                            const double* axz, const double* T, const double* u, const double* v, const double* rho);
void      orc_atmo_destroy(orc_atmo* a);
void      orc_atmo_sample(orc_atmo* a, double p0, double p1, double p2, double* out4);   /* c, u, v, rho */
int64_t   orc_atmo3d_slopes(const orc_atmo* at, int field, int which, double* out);   /* Set_Slopes_Multi output, see orc_mspline.c */

double orc_suthbass_alpha(orc_atmo* a, double x0, double x1, double x2, double freq);

/* Mirror of Load_G2S (Code/Atmo/G2S_Spline1D.cpp:109-142, G2S_GlobalSpline1D.cpp:113-152). */
int orc_load_met_1d(const char* path, const char* format, double z_grnd_taper, int global_taper,
                    int cap, int* n, double* z, double* T, double* u, double* v, double* rho);

/* Mirror of Load_G2S_Multi (Code/Atmo/G2S_MultiDimSpline3D.cpp:139-189, G2S_GlobalMultiDimSpline3D.cpp:142-199). */
int orc_load_met_grid(const char* prefix, const char* loc0, const char* loc1, const char* format, int global,
                      int cap0, int cap1, int capz, int* n0, int* n1, int* nz,
                      double* ax0, double* ax1, double* axz, double* T, double* u, double* v, double* rho);

/* Same contract as geoac_trace() of include/geoac_b200.h but on the CPU, single thread.
 * `limits_from_atmo` != 0 applies GeoAc_SetPropRegion to a copy of *p first. Returns total RK4 steps (<0 on error). */
int64_t orc_trace(int variant, orc_atmo* atmo, const geoac_params* p, int64_t n_rays,
                  const double* theta, const double* phi, double* rec, int32_t* status, int32_t* n_steps);
/* orc_trace plus the raypath rows of WriteRays=True (one row of GEOAC_PATH_NF doubles every path_stride steps, at most
 * path_cap rows kept per ray, path_rows[ray] = rows produced); needs accum_per_segment semantics. */
int64_t orc_trace_paths(int variant, orc_atmo* atmo, const geoac_params* p, int64_t n_rays,
                        const double* theta, const double* phi, double* rec, int32_t* status, int32_t* n_steps,
                        int path_stride, int64_t path_cap, double* path, int32_t* path_rows,
                        int64_t caus_cap, double* caus, int32_t* caus_rows);

/* GeoAc_SetPropRegion for this atmosphere: fills vert_limit / range_limit / box limits of *p. */
void orc_set_prop_region(int variant, const orc_atmo* atmo, geoac_params* p);

#ifdef __cplusplus
}
#endif
#endif
