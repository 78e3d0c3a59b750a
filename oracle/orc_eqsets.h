/* oracle/orc_eqsets.h -- TEST INFRASTRUCTURE (CPU oracle). Internal: per-variant equation-set dispatch table. */
#ifndef ORC_EQSETS_H_
#define ORC_EQSETS_H_
#include "geoac_oracle.h"

/* everything the reference keeps in globals (GeoAc_theta, GeoAc_phi, GeoAc_Sources, limits) for ONE ray */
typedef struct orc_ray {
    orc_atmo* atmo;
    const geoac_params* prm;
    double theta, phi;          /* GeoAc_theta, GeoAc_phi [rad] */
    int eq_cnt, calc_amp;
    double S[160];              /* variant-specific "GeoAc_Sources" scratch, laid out by each equation set */
} orc_ray;

typedef struct orc_eqset {
    int eq_amp, eq_noamp;
    void   (*init)(orc_ray*, double* y0);
    void   (*update)(orc_ray*, const double* y);
    double (*rhs)(orc_ray*, const double* y, int i);
    double (*set_ds)(orc_ray*, const double* y);
    int    (*brk)(orc_ray*, const double* y);
    int    (*gnd)(orc_ray*, const double* y);
    void   (*tt_seg)(orc_ray*, const double* ya, const double* yb, double* acc);
    void   (*sb_seg)(orc_ray*, const double* ya, const double* yb, double* acc);
    double (*amplitude)(orc_ray*, const double* yk);
    double (*jacobian)(orc_ray*, const double* yk);      /* GeoAc_Jacobian: sign changes mark caustics */
    void   (*reflect)(orc_ray*, const double* ykm2, const double* ykm1, const double* yk, double* y0);
    double (*altitude)(orc_ray*, const double* y);
    /* fills inclination / back azimuth / aux / margin of the record (theta_deg-free: uses ray->theta/phi) */
    void   (*finish)(orc_ray*, const double* ykm1, const double* yk, double travel_time, double* incl, double* backaz,
                     double* aux, double* margin);
} orc_eqset;

extern const orc_eqset orc_eq_2d, orc_eq_3d, orc_eq_global, orc_eq_3drngdep, orc_eq_globalrngdep;

#define ORC_PI 3.141592653589793238462643

#endif
