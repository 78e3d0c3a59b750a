#include "orc_eqsets.h"
const orc_eqset orc_eq_global = {0}, orc_eq_3drngdep = {0}, orc_eq_globalrngdep = {0};
void orc_mspline_free(void* g) {(void)g;}
void orc_mspline_region(const orc_atmo* a, geoac_params* p) {(void)a;(void)p;}
orc_atmo* orc_atmo3d_create(int global, int n0, int n1, int nz, const double* ax0, const double* ax1,
                            const double* axz, const double* T, const double* u, const double* v, const double* rho) { return 0; }
