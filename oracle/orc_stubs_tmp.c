/* temporary: range-dependent pieces not restated yet */
#include "orc_eqsets.h"
const orc_eqset orc_eq_3drngdep = {0};
void orc_mspline_free(void* g) {(void)g;}
void orc_mspline_region(const orc_atmo* a, geoac_params* p) {(void)a;(void)p;}
void orc_mspline_allorder1(orc_atmo* a, int field, double q0, double q1, double q2, double* f, double d[3]) {}
void orc_mspline_allorder2(orc_atmo* a, int field, double q0, double q1, double q2, double* f, double d[3], double dd[3][3]) {}
orc_atmo* orc_atmo3d_create(int global, int n0, int n1, int nz, const double* ax0, const double* ax1,
                            const double* axz, const double* T, const double* u, const double* v, const double* rho) { return 0; }
