/* oracle/orc_mspline.c -- TEST INFRASTRUCTURE (CPU oracle), never linked into or called by the product.
 *
 * Range-dependent atmosphere: vertical natural cubic splines at every horizontal node x bicubic Hermite patches built
 * from finite differences of vertical-spline values.  Restates, with the same floating-point expression trees,
 *   Cartesian: Code/Atmo/G2S_MultiDimSpline3D.cpp       (Set_Slopes_Multi :306-425, Find_Segment :432-474,
 *              Eval_Vert_Spline_* :477-562, BiCubic_Deriv_* :568-800, Eval_Spline_f/df :806-971,
 *              Eval_Spline_AllOrder1/2 :1156-1593, wrappers :1633-1743, Load_G2S_Multi :139-189)
 *   Global   : Code/Atmo/G2S_GlobalMultiDimSpline3D.cpp (same roles at :313-431, :432-474, :477-565, :571-755,
 *              :757-880, :1047-1461, :1502-1611, :142-199)
 * The two reference files are near-copies with different slips; one parameterised
 * implementation with `g->global` selecting the slip keeps them side by side.  Arrays are dense [n0][n1][nz]
 * (vertical index fastest) for both -- the reference's [r][t][p] pointer layout does not change any arithmetic.
 * Eval_Spline_ddf (c_ddiff, u_ddiff, v_ddiff) is not on the -prop path and is not restated.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "geoac_oracle.h"

typedef struct ms_field {
    double *f, *s, *sa, *sb;        /* values, vertical slopes of f, of df/d(ax0), of df/d(ax1) */
    int accel[3];                   /* cursors: ax0, ax1, vertical (Spline.accel) */
} ms_field;

typedef struct ms_grid {
    int global, n0, n1, nz;
    double *a, *b, *z;              /* node coordinates: ax0 (x | lat), ax1 (y | lon), vertical (z | r) */
    double amin, amax, bmin, bmax, zmin, zmax;
    ms_field F[4];                  /* 0 T, 1 u, 2 v, 3 rho */
} ms_grid;

#define IDX(g, i, j, k) ((((size_t)(i)) * (g)->n1 + (j)) * (g)->nz + (k))

static const double M16[16][16] = {          /* BiCubic_ConversionMatrix, G2S_MultiDimSpline3D.cpp:213-230 */
    { 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, { 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
    {-3, 3, 0, 0,-2,-1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, { 2,-2, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
    { 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0}, { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0},
    { 0, 0, 0, 0, 0, 0, 0, 0,-3, 3, 0, 0,-2,-1, 0, 0}, { 0, 0, 0, 0, 0, 0, 0, 0, 2,-2, 0, 0, 1, 1, 0, 0},
    {-3, 0, 3, 0, 0, 0, 0, 0,-2, 0,-1, 0, 0, 0, 0, 0}, { 0, 0, 0, 0,-3, 0, 3, 0, 0, 0, 0, 0,-2, 0,-1, 0},
    { 9,-9,-9, 9, 6, 3,-6,-3, 6,-6, 3,-3, 4, 2, 2, 1}, {-6, 6, 6,-6,-3,-3, 3, 3,-4, 4,-2, 2,-2,-2,-1,-1},
    { 2, 0,-2, 0, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 0, 0}, { 0, 0, 0, 0, 2, 0,-2, 0, 0, 0, 0, 0, 1, 0, 1, 0},
    {-6, 6, 6,-6,-4,-2, 4, 2,-3, 3,-3, 3,-2,-1,-2,-1}, { 4,-4,-4, 4, 2, 2,-2,-2, 2,-2, 2,-2, 1, 1, 1, 1}
};

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* ---- Set_Slopes_Multi: Thomas solve per column for f, then for the node finite differences d/d(ax0), d/d(ax1) ---- */
static void column_slopes(const ms_grid* g, const double* vals /* [nz] */, double* out /* [nz] */, int shifted_diff) {
    const int n = g->nz; const double* z = g->z;
    double* nc = (double*)malloc(sizeof(double) * (size_t)n);
    double* nd = (double*)malloc(sizeof(double) * (size_t)n);
    double ai, bi, ci, di;
    bi = 2.0 / (z[1] - z[0]); ci = 1.0 / (z[1] - z[0]);
    di = 3.0 * (vals[1] - vals[0]) / pow(z[1] - z[0], 2);
    nc[0] = ci / bi; nd[0] = di / bi;
    for (int i = 1; i < n - 1; i++) {
        ai = 1.0 / (z[i] - z[i - 1]);
        bi = 2.0 * (1.0 / (z[i] - z[i - 1]) + 1.0 / (z[i + 1] - z[i]));
        ci = 1.0 / (z[i + 1] - z[i]);
        if (shifted_diff)   /* Global dfdt/dfdp slopes: dfdt[i] - dfdt[i+1] in the first term (App. A-9, Global :381,414) */
            di = 3.0 * ((vals[i] - vals[i + 1]) / pow(z[i] - z[i - 1], 2) + (vals[i + 1] - vals[i]) / pow(z[i + 1] - z[i], 2));
        else
            di = 3.0 * ((vals[i] - vals[i - 1]) / pow(z[i] - z[i - 1], 2) + (vals[i + 1] - vals[i]) / pow(z[i + 1] - z[i], 2));
        nc[i] = ci / (bi - nc[i - 1] * ai);
        nd[i] = (di - nd[i - 1] * ai) / (bi - nc[i - 1] * ai);
    }
    ai = 1.0 / (z[n - 1] - z[n - 2]); bi = 2.0 / (z[n - 1] - z[n - 2]);
    di = 3.0 * (vals[n - 1] - vals[n - 2]) / pow(z[n - 1] - z[n - 2], 2);
    nd[n - 1] = (di - nd[n - 2] * ai) / (bi - nc[n - 2] * ai);
    out[n - 1] = nd[n - 1];
    for (int i = n - 2; i >= 0; i--) out[i] = nd[i] - nc[i] * out[i + 1];
    free(nc); free(nd);
}

static void set_slopes_multi(const ms_grid* g, ms_field* F) {
    const int nz = g->nz;
    double* da = (double*)malloc(sizeof(double) * (size_t)nz);
    double* db = (double*)malloc(sizeof(double) * (size_t)nz);
    for (int i = 0; i < g->n0; i++) for (int j = 0; j < g->n1; j++) {
        column_slopes(g, F->f + IDX(g, i, j, 0), F->s + IDX(g, i, j, 0), 0);
        int iu = imin(i + 1, g->n0 - 1), id = imax(i - 1, 0), ju = imin(j + 1, g->n1 - 1), jd = imax(j - 1, 0);
        for (int k = 0; k < nz; k++) {
            da[k] = (F->f[IDX(g, iu, j, k)] - F->f[IDX(g, id, j, k)]) / (g->a[iu] - g->a[id]);
            db[k] = (F->f[IDX(g, i, ju, k)] - F->f[IDX(g, i, jd, k)]) / (g->b[ju] - g->b[jd]);
        }
        column_slopes(g, da, F->sa + IDX(g, i, j, 0), g->global);
        column_slopes(g, db, F->sb + IDX(g, i, j, 0), g->global);
    }
    free(da); free(db);
}

/* ---- Find_Segment (both files) ---- */
static int find_segment(double x, const double* xv, int length, int* prev) {
    int index = length + 1, done = 0;
    if (x >= xv[*prev] && x <= xv[*prev + 1]) done = 1;
    if (!done && *prev + 2 <= length - 1) { if (x >= xv[*prev + 1] && x <= xv[*prev + 2]) { done = 1; *prev = *prev + 1; } }
    if (!done && *prev - 1 >= 0)          { if (x >= xv[*prev - 1] && x <= xv[*prev])     { done = 1; *prev = *prev - 1; } }
    if (!done) {
        for (int i = 0; i < length; i++) {
            if (x >= xv[i] && x <= xv[i + 1]) { index = i; break; }
            if (x >= xv[length - 2 - i] && x < xv[length - 1 - i]) { index = (length - 2) - i; break; }
        }
        *prev = index;
    }
    return *prev;
}

/* ---- column primitives: Eval_Vert_Spline_* ---- */
typedef struct qry { const ms_grid* g; const ms_field* F; double z; int kz; } qry;

static double V(const qry* q, int i, int j) {                       /* Eval_Vert_Spline_f */
    const ms_grid* g = q->g; const double* zv = g->z; int kz = q->kz; double z = q->z;
    const double* f = q->F->f + IDX(g, i, j, 0); const double* s = q->F->s + IDX(g, i, j, 0);
    double X = (z - zv[kz]) / (zv[kz + 1] - zv[kz]);
    double A = s[kz] * (zv[kz + 1] - zv[kz]) - (f[kz + 1] - f[kz]);
    double B = -s[kz + 1] * (zv[kz + 1] - zv[kz]) + (f[kz + 1] - f[kz]);
    return (1.0 - X) * f[kz] + X * f[kz + 1] + X * (1.0 - X) * (A * (1.0 - X) + B * X);
}
static double Vz(const qry* q, int i, int j) {                      /* Eval_Vert_Spline_dfdz | dfdr */
    const ms_grid* g = q->g; const double* zv = g->z; int kz = q->kz; double z = q->z;
    const double* f = q->F->f + IDX(g, i, j, 0); const double* s = q->F->s + IDX(g, i, j, 0);
    double X = (z - zv[kz]) / (zv[kz + 1] - zv[kz]);
    double A = s[kz] * (zv[kz + 1] - zv[kz]) - (f[kz + 1] - f[kz]);
    double B = -s[kz + 1] * (zv[kz + 1] - zv[kz]) + (f[kz + 1] - f[kz]);
    return (f[kz + 1] - f[kz]) / (zv[kz + 1] - zv[kz])
         + (1.0 - 2.0 * X) * (A * (1.0 - X) + B * X) / (zv[kz + 1] - zv[kz])
         + X * (1.0 - X) * (B - A) / (zv[kz + 1] - zv[kz]);
}
static double Vzz(const qry* q, int i, int j) {                     /* Eval_Vert_Spline_ddfdzdz | ddfdrdr */
    const ms_grid* g = q->g; const double* zv = g->z; int kz = q->kz; double z = q->z;
    const double* f = q->F->f + IDX(g, i, j, 0); const double* s = q->F->s + IDX(g, i, j, 0);
    double X = (z - zv[kz]) / (zv[kz + 1] - zv[kz]);
    double A = s[kz] * (zv[kz + 1] - zv[kz]) - (f[kz + 1] - f[kz]);
    double B = -s[kz + 1] * (zv[kz + 1] - zv[kz]) + (f[kz + 1] - f[kz]);
    return 2.0 * (B - 2.0 * A + (A - B) * 3.0 * X) / pow(zv[kz + 1] - zv[kz], 2);
}
/* vertical spline of the node finite difference along axis `ax` (0: ax0, 1: ax1); deriv = 0 value, 1 vertical derivative */
static double G(const qry* q, int i, int j, int ax, int deriv) {    /* Eval_Vert_Spline_dfdx|dfdy|ddfdxdz|ddfdydz (+ Global names) */
    const ms_grid* g = q->g; const double* zv = g->z; int kz = q->kz; double z = q->z;
    const ms_field* F = q->F;
    double d0, d1; const double* sl;
    if (ax == 0) {
        int up = imin(i + 1, g->n0 - 1), dn = imax(i - 1, 0);
        d0 = (F->f[IDX(g, up, j, kz)] - F->f[IDX(g, dn, j, kz)]) / (g->a[up] - g->a[dn]);
        d1 = (F->f[IDX(g, up, j, kz + 1)] - F->f[IDX(g, dn, j, kz + 1)]) / (g->a[up] - g->a[dn]);
        sl = F->sa + IDX(g, i, j, 0);
    } else {
        int up = imin(j + 1, g->n1 - 1), dn = imax(j - 1, 0);
        d0 = (F->f[IDX(g, i, up, kz)] - F->f[IDX(g, i, dn, kz)]) / (g->b[up] - g->b[dn]);
        d1 = (F->f[IDX(g, i, up, kz + 1)] - F->f[IDX(g, i, dn, kz + 1)]) / (g->b[up] - g->b[dn]);
        sl = F->sb + IDX(g, i, j, 0);
    }
    double X = (z - zv[kz]) / (zv[kz + 1] - zv[kz]);
    double A = sl[kz] * (zv[kz + 1] - zv[kz]) - (d1 - d0);
    double B = -sl[kz + 1] * (zv[kz + 1] - zv[kz]) + (d1 - d0);
    if (!deriv) return (1.0 - X) * d0 + X * d1 + X * (1.0 - X) * (A * (1.0 - X) + B * X);
    double lead = g->global ? (d1 - d1) : (d1 - d0);            /* Global: (dfdt_krp1 - dfdt_krp1), App. A-9 */
    return lead / (zv[kz + 1] - zv[kz])
         + (1.0 - 2.0 * X) * (A * (1.0 - X) + B * X) / (zv[kz + 1] - zv[kz])
         + X * (1.0 - X) * (B - A) / (zv[kz + 1] - zv[kz]);
}

/* column kinds the finite-difference operators act on */
enum { C_V, C_VZ, C_VZZ, C_GA, C_GB };
static double col(const qry* q, int kind, int i, int j) {
    switch (kind) {
        case C_V: return V(q, i, j);
        case C_VZ: return Vz(q, i, j);
        case C_VZZ: return Vzz(q, i, j);
        case C_GA: return G(q, i, j, 0, 0);
        default: return G(q, i, j, 1, 0);
    }
}
/* BiCubic_Deriv_*: centred differences of a column quantity, one-sided at the grid edge */
static double fd_a(const qry* q, int kind, int i, int j) {
    int up = i + 1, dn = i - 1;
    if (up > q->g->n0 - 1) up = i;
    if (dn < 0) dn = i;
    return (col(q, kind, up, j) - col(q, kind, dn, j)) / (q->g->a[up] - q->g->a[dn]);
}
static double fd_b(const qry* q, int kind, int i, int j) {
    int up = j + 1, dn = j - 1;
    if (up > q->g->n1 - 1) up = j;
    if (dn < 0) dn = j;
    return (col(q, kind, i, up) - col(q, kind, i, dn)) / (q->g->b[up] - q->g->b[dn]);
}
static double fd_ab(const qry* q, int kind, int i, int j) {
    int iu = i + 1, id = i - 1, ju = j + 1, jd = j - 1;
    if (iu > q->g->n0 - 1) iu = i;
    if (id < 0) id = i;
    if (ju > q->g->n1 - 1) ju = j;
    if (jd < 0) jd = j;
    return (col(q, kind, iu, ju) - col(q, kind, iu, jd) - col(q, kind, id, ju) + col(q, kind, id, jd))
         / ((q->g->a[iu] - q->g->a[id]) * (q->g->b[ju] - q->g->b[jd]));
}

static void mat16(const double* X, double* A) {
    for (int j = 0; j < 16; j++) { A[j] = 0; for (int k = 0; k < 16; k++) A[j] += M16[j][k] * X[k]; }
}
static double poly(const double* A, double xs, double ys) {
    double r = 0;
    for (int k1 = 0; k1 < 4; k1++) for (int k2 = 0; k2 < 4; k2++) r += 1.0 * A[k1 + 4 * k2] * pow(xs, k1) * pow(ys, k2);
    return r;
}
static double poly_da(const double* A, double xs, double ys) {      /* d/d(xs) */
    double r = 0;
    for (int k1 = 1; k1 < 4; k1++) for (int k2 = 0; k2 < 4; k2++) r += 1.0 * k1 * A[k1 + 4 * k2] * pow(xs, k1 - 1) * pow(ys, k2);
    return r;
}
static double poly_db(const double* A, double xs, double ys) {      /* d/d(ys) */
    double r = 0;
    for (int k1 = 0; k1 < 4; k1++) for (int k2 = 1; k2 < 4; k2++) r += 1.0 * k2 * A[k1 + 4 * k2] * pow(xs, k1) * pow(ys, k2 - 1);
    return r;
}
static double poly_da_div(const double* A, double xs, double ys, double d) {   /* Cartesian: each term divided by dx */
    double r = 0;
    for (int k1 = 1; k1 < 4; k1++) for (int k2 = 0; k2 < 4; k2++) r += 1.0 * k1 * A[k1 + 4 * k2] * pow(xs, k1 - 1) * pow(ys, k2) / d;
    return r;
}
static double poly_db_div(const double* A, double xs, double ys, double d) {
    double r = 0;
    for (int k1 = 0; k1 < 4; k1++) for (int k2 = 1; k2 < 4; k2++) r += 1.0 * k2 * A[k1 + 4 * k2] * pow(xs, k1) * pow(ys, k2 - 1) / d;
    return r;
}

/* corner order of X_vec blocks: (i,j), (i+1,j), (i,j+1), (i+1,j+1) */
static const int CI[4] = { 0, 1, 0, 1 }, CJ[4] = { 0, 0, 1, 1 };

typedef struct cell { qry q; int ka, kb; double da, db, xs, ys; } cell;

static void locate(ms_grid* g, ms_field* F, double a, double b, double z, cell* c) {
    c->q.g = g; c->q.F = F; c->q.z = z;
    if (g->global) {            /* Global: Find_Segment order r, t, p with accel[0..2] = r, t, p */
        c->q.kz = find_segment(z, g->z, g->nz, &F->accel[2]);
        c->ka = find_segment(a, g->a, g->n0, &F->accel[0]);
        c->kb = find_segment(b, g->b, g->n1, &F->accel[1]);
    } else {
        c->ka = find_segment(a, g->a, g->n0, &F->accel[0]);
        c->kb = find_segment(b, g->b, g->n1, &F->accel[1]);
        c->q.kz = find_segment(z, g->z, g->nz, &F->accel[2]);
    }
    c->da = g->a[c->ka + 1] - g->a[c->ka];
    c->db = g->b[c->kb + 1] - g->b[c->kb];
    c->xs = (a - g->a[c->ka]) / (g->a[c->ka + 1] - g->a[c->ka]);
    c->ys = (b - g->b[c->kb]) / (g->b[c->kb + 1] - g->b[c->kb]);
}

/* Eval_Spline_f: G2S_MultiDimSpline3D.cpp:806-866 (y data scaled by dx_scalar, App. A-8) / Global :757-815 */
static double eval_f(ms_grid* g, ms_field* F, double a, double b, double z) {
    cell c; locate(g, F, a, b, z, &c);
    double X[16], A[16];
    const double sb = g->global ? c.db : c.da;
    for (int m = 0; m < 4; m++) {
        int i = c.ka + CI[m], j = c.kb + CJ[m];
        X[m] = V(&c.q, i, j);
        X[4 + m] = fd_a(&c.q, C_V, i, j) * c.da;
        X[8 + m] = fd_b(&c.q, C_V, i, j) * sb;
        X[12 + m] = fd_ab(&c.q, C_V, i, j) * c.da * c.db;
    }
    mat16(X, A);
    return poly(A, c.xs, c.ys);
}

/* Eval_Spline_df: `axis` 0 = ax0, 1 = ax1, 2 = vertical (the callers map the reference's index convention) */
static double eval_df(ms_grid* g, ms_field* F, double a, double b, double z, int axis) {
    cell c; locate(g, F, a, b, z, &c);
    double X[16], A[16];
    const double sb = g->global ? c.db : c.da;
    for (int m = 0; m < 4; m++) {
        int i = c.ka + CI[m], j = c.kb + CJ[m];
        if (axis == 0) {
            X[m] = fd_a(&c.q, C_V, i, j);
            X[4 + m] = fd_a(&c.q, C_GA, i, j) * c.da;
            X[8 + m] = fd_ab(&c.q, C_V, i, j) * sb;
            X[12 + m] = fd_ab(&c.q, C_GA, i, j) * c.da * c.db;
        } else if (axis == 1) {
            X[m] = fd_b(&c.q, C_V, i, j);
            X[4 + m] = fd_ab(&c.q, C_V, i, j) * c.da;
            X[8 + m] = fd_b(&c.q, C_GB, i, j) * sb;
            X[12 + m] = fd_ab(&c.q, C_GB, i, j) * c.da * c.db;
        } else {
            X[m] = Vz(&c.q, i, j);
            X[4 + m] = G(&c.q, i, j, 0, 1) * c.da;
            X[8 + m] = G(&c.q, i, j, 1, 1) * sb;
            X[12 + m] = fd_ab(&c.q, C_VZ, i, j) * c.da * c.db;
        }
    }
    mat16(X, A);
    return poly(A, c.xs, c.ys);
}

static double clampd(double v, double lo, double hi) { double e = fmin(v, hi); e = fmax(e, lo); return e; }

/* Eval_Spline_AllOrder1 / AllOrder2.  Outputs in grid-axis order: d[0..2] = d/d(ax0), d/d(ax1), d/d(vertical);
 * dd[n][m] likewise (symmetric).  order2 == 0 leaves dd untouched. */
static void allorder(ms_grid* g, ms_field* F, double a_in, double b_in, double z_in, int order2, double* f, double d[3], double dd[3][3]) {
    double a = clampd(a_in, g->amin, g->amax), b = clampd(b_in, g->bmin, g->bmax), z = clampd(z_in, g->zmin, g->zmax);
    cell c; locate(g, F, a, b, z, &c);
    double X[16], A[16];
    double Fa[4], Fb[4], Fab[4];
    int I[4], J[4];
    for (int m = 0; m < 4; m++) { I[m] = c.ka + CI[m]; J[m] = c.kb + CJ[m]; }
    for (int m = 0; m < 4; m++) Fa[m] = fd_a(&c.q, C_V, I[m], J[m]);
    for (int m = 0; m < 4; m++) Fb[m] = fd_b(&c.q, C_V, I[m], J[m]);
    for (int m = 0; m < 4; m++) Fab[m] = fd_ab(&c.q, C_V, I[m], J[m]);
    const int glob = g->global;

    /* fit F */
    for (int m = 0; m < 4; m++) { X[m] = V(&c.q, I[m], J[m]); X[4 + m] = Fa[m] * c.da; X[8 + m] = Fb[m] * c.db; X[12 + m] = Fab[m] * c.da * c.db; }
    mat16(X, A);
    *f = poly(A, c.xs, c.ys);

    if (glob) {   /* Global evaluates the vertical-derivative fit second (order only matters for reading along) */
        for (int m = 0; m < 4; m++) {
            X[m] = Vz(&c.q, I[m], J[m]); X[4 + m] = G(&c.q, I[m], J[m], 0, 1) * c.da; X[8 + m] = G(&c.q, I[m], J[m], 1, 1) * c.db;
            X[12 + m] = fd_ab(&c.q, C_VZ, I[m], J[m]) * c.da * c.db;
        }
        mat16(X, A);
        d[2] = poly(A, c.xs, c.ys);
        if (order2) { dd[2][0] = dd[0][2] = poly_da(A, c.xs, c.ys); dd[2][1] = dd[1][2] = poly_db(A, c.xs, c.ys); }   /* no /dt, /dp: App. A-9 */
    }

    /* fit FA (d/d ax0) */
    for (int m = 0; m < 4; m++) {
        X[m] = Fa[m]; X[4 + m] = fd_a(&c.q, C_GA, I[m], J[m]) * c.da; X[8 + m] = Fab[m] * c.db;
        X[12 + m] = fd_ab(&c.q, C_GA, I[m], J[m]) * c.da * c.db;
    }
    mat16(X, A);
    d[0] = poly(A, c.xs, c.ys);
    if (order2) {
        if (glob) { dd[0][0] = poly_da(A, c.xs, c.ys); dd[0][1] = dd[1][0] = poly_db(A, c.xs, c.ys); }
        else      { dd[0][0] = poly_da_div(A, c.xs, c.ys, c.da); dd[0][1] = dd[1][0] = poly_db_div(A, c.xs, c.ys, c.db); }
    }

    /* fit FB (d/d ax1) */
    for (int m = 0; m < 4; m++) {
        X[m] = Fb[m]; X[4 + m] = Fab[m] * c.da; X[8 + m] = fd_b(&c.q, C_GB, I[m], J[m]) * c.db;
        X[12 + m] = fd_ab(&c.q, C_GB, I[m], J[m]) * c.da * c.db;
    }
    mat16(X, A);
    d[1] = poly(A, c.xs, c.ys);
    if (order2) dd[1][1] = glob ? poly_db(A, c.xs, c.ys) : poly_db_div(A, c.xs, c.ys, c.db);

    if (!glob) {  /* fit FZ */
        for (int m = 0; m < 4; m++) {
            X[m] = Vz(&c.q, I[m], J[m]); X[4 + m] = G(&c.q, I[m], J[m], 0, 1) * c.da; X[8 + m] = G(&c.q, I[m], J[m], 1, 1) * c.db;
            X[12 + m] = fd_ab(&c.q, C_VZ, I[m], J[m]) * c.da * c.db;
        }
        mat16(X, A);
        d[2] = poly(A, c.xs, c.ys);
        if (order2) { dd[0][2] = dd[2][0] = poly_da_div(A, c.xs, c.ys, c.da); dd[1][2] = dd[2][1] = poly_db_div(A, c.xs, c.ys, c.db); }
    }

    if (order2) { /* fit FZZ: Cartesian scales the ax1 data by dx_scalar (App. A-8, :1568-1571) */
        const double sb = glob ? c.db : c.da;
        for (int m = 0; m < 4; m++) {
            X[m] = Vzz(&c.q, I[m], J[m]); X[4 + m] = fd_a(&c.q, C_VZZ, I[m], J[m]) * c.da; X[8 + m] = fd_b(&c.q, C_VZZ, I[m], J[m]) * sb;
            X[12 + m] = fd_ab(&c.q, C_VZZ, I[m], J[m]) * c.da * c.db;
        }
        mat16(X, A);
        dd[2][2] = poly(A, c.xs, c.ys);
    }
}

/* ------------------------------------------------------------------------------------------------------------------
 * Atmosphere API on top (Atmo_State.h): coordinates arrive in the variant's own order -- Cartesian (x, y, z),
 * Global (r, lat, lon) -- and are mapped to grid axes (ax0, ax1, vertical).
 * ------------------------------------------------------------------------------------------------------------------ */
static ms_grid* GRID(orc_atmo* a) { return (ms_grid*)a->grid; }
static void to_axes(const ms_grid* g, double p0, double p1, double p2, double* a, double* b, double* z) {
    if (g->global) { *z = p0; *a = p1; *b = p2; } else { *a = p0; *b = p1; *z = p2; }
}
static int axis_of_index(const ms_grid* g, int n) { return g->global ? (n == 0 ? 2 : n - 1) : n; }

static double wf(orc_atmo* at, int field, double p0, double p1, double p2) {
    ms_grid* g = GRID(at); double a, b, z; to_axes(g, p0, p1, p2, &a, &b, &z);
    return eval_f(g, &g->F[field], clampd(a, g->amin, g->amax), clampd(b, g->bmin, g->bmax), clampd(z, g->zmin, g->zmax));
}
static double wdf(orc_atmo* at, int field, double p0, double p1, double p2, int n) {
    ms_grid* g = GRID(at); double a, b, z; to_axes(g, p0, p1, p2, &a, &b, &z);
    return eval_df(g, &g->F[field], clampd(a, g->amin, g->amax), clampd(b, g->bmin, g->bmax), clampd(z, g->zmin, g->zmax), axis_of_index(g, n));
}
static const double GAMR = 0.00040187;
static double m_rho(orc_atmo* a, double p0, double p1, double p2) { return wf(a, 3, p0, p1, p2); }
static double m_c(orc_atmo* a, double p0, double p1, double p2) { return sqrt(GAMR * wf(a, 0, p0, p1, p2)); }
static double m_c_diff(orc_atmo* a, double p0, double p1, double p2, int n) { return GAMR / (2.0 * m_c(a, p0, p1, p2)) * wdf(a, 0, p0, p1, p2, n); }
static double m_u(orc_atmo* a, double p0, double p1, double p2) { return wf(a, 1, p0, p1, p2); }
static double m_u_diff(orc_atmo* a, double p0, double p1, double p2, int n) { return wdf(a, 1, p0, p1, p2, n); }
static double m_v(orc_atmo* a, double p0, double p1, double p2) { return wf(a, 2, p0, p1, p2); }
static double m_v_diff(orc_atmo* a, double p0, double p1, double p2, int n) { return wdf(a, 2, p0, p1, p2, n); }
static double m_dd_unused(orc_atmo* a, double p0, double p1, double p2, int n1, int n2) {
    (void)a; (void)p0; (void)p1; (void)p2; (void)n1; (void)n2; return NAN;    /* Eval_Spline_ddf: not on the -prop path */
}

/* fast paths used by the range-dependent GeoAc_UpdateSources; outputs in the VARIANT's coordinate order */
static void reorder(const ms_grid* g, const double d[3], double dd[3][3], int order2, double dout[3], double ddout[3][3]) {
    /* grid axes (ax0, ax1, vert) -> Cartesian (x, y, z) identity; Global (r, t, p) = (vert, ax0, ax1) */
    int map[3]; if (g->global) { map[0] = 2; map[1] = 0; map[2] = 1; } else { map[0] = 0; map[1] = 1; map[2] = 2; }
    for (int n = 0; n < 3; n++) dout[n] = d[map[n]];
    if (order2) for (int n = 0; n < 3; n++) for (int m = 0; m < 3; m++) ddout[n][m] = dd[map[n]][map[m]];
}
void orc_mspline_allorder1(orc_atmo* at, int field, double q0, double q1, double q2, double* f, double dout[3]) {
    ms_grid* g = GRID(at); double a, b, z; to_axes(g, q0, q1, q2, &a, &b, &z);
    double d[3], dd[3][3];
    allorder(g, &g->F[field], a, b, z, 0, f, d, dd);
    reorder(g, d, dd, 0, dout, dd);
}
void orc_mspline_allorder2(orc_atmo* at, int field, double q0, double q1, double q2, double* f, double dout[3], double ddout[3][3]) {
    ms_grid* g = GRID(at); double a, b, z; to_axes(g, q0, q1, q2, &a, &b, &z);
    double d[3], dd[3][3];
    allorder(g, &g->F[field], a, b, z, 1, f, d, dd);
    reorder(g, d, dd, 1, dout, ddout);
}
/* Windu/Windv cursors follow Temp's after the first evaluation of a stage (3DRngDep.cpp:227-234) */
void orc_mspline_sync_accel(orc_atmo* at) {
    ms_grid* g = GRID(at);
    for (int n = 0; n < 3; n++) { g->F[1].accel[n] = g->F[0].accel[n]; g->F[2].accel[n] = g->F[0].accel[n]; }
}
/* GeoAc_SetInitialConditions resets the Temp/Windu/Windv cursors (3DRngDep.cpp:130-134) */
void orc_mspline_reset_accel(orc_atmo* at) {
    ms_grid* g = GRID(at);
    for (int f = 0; f < 3; f++) for (int n = 0; n < 3; n++) g->F[f].accel[n] = 0;
}

void orc_mspline_region(const orc_atmo* at, geoac_params* p) {      /* GeoAc_SetPropRegion */
    const ms_grid* g = (const ms_grid*)at->grid;
    p->vert_limit = g->zmax;
    p->box_min[0] = g->amin; p->box_max[0] = g->amax; p->box_min[1] = g->bmin; p->box_max[1] = g->bmax;
}

void orc_mspline_free(void* grid) {
    ms_grid* g = (ms_grid*)grid;
    if (!g) return;
    for (int f = 0; f < 4; f++) { free(g->F[f].f); free(g->F[f].s); free(g->F[f].sa); free(g->F[f].sb); }
    free(g->a); free(g->b); free(g->z); free(g);
}

static double* dupn(const double* s, size_t n) { double* d = (double*)malloc(sizeof(double) * n); memcpy(d, s, sizeof(double) * n); return d; }

/* Spline_Multi_G2S minus file I/O.  ax0/ax1: x,y [km] or lat,lon [rad]; axz: altitude [km above sea level];
 * fields dense [n0][n1][nz], winds already in km/s and tapered. */
orc_atmo* orc_atmo3d_create(int global, int n0, int n1, int nz, const double* ax0, const double* ax1,
                            const double* axz, const double* T, const double* u, const double* v, const double* rho) {
    if (n0 < 2 || n1 < 2 || nz < 3) return 0;
    orc_atmo* at = (orc_atmo*)calloc(1, sizeof(orc_atmo));
    ms_grid* g = (ms_grid*)calloc(1, sizeof(ms_grid));
    const size_t N = (size_t)n0 * n1 * nz;
    g->global = global; g->n0 = n0; g->n1 = n1; g->nz = nz;
    g->a = dupn(ax0, n0); g->b = dupn(ax1, n1); g->z = dupn(axz, nz);
    if (global) for (int k = 0; k < nz; k++) g->z[k] += 6370.0;                 /* r_vals[nr] += r_earth */
    g->amin = g->a[0]; g->amax = g->a[n0 - 1]; g->bmin = g->b[0]; g->bmax = g->b[n1 - 1]; g->zmin = g->z[0]; g->zmax = g->z[nz - 1];
    const double* src[4] = { T, u, v, rho };
    for (int f = 0; f < 4; f++) {
        g->F[f].f = dupn(src[f], N);
        g->F[f].s = (double*)malloc(sizeof(double) * N); g->F[f].sa = (double*)malloc(sizeof(double) * N); g->F[f].sb = (double*)malloc(sizeof(double) * N);
        set_slopes_multi(g, &g->F[f]);
    }
    at->kind = global ? 3 : 2; at->vert_index = global ? 0 : 2;
    at->vmin = g->zmin; at->vmax = g->zmax; at->r_earth = 6370.0; at->z_grnd = 0.0; at->tweak_abs = 0.3;
    at->grid = g;
    at->c = m_c; at->c_diff = m_c_diff; at->c_ddiff = m_dd_unused;
    at->u = m_u; at->u_diff = m_u_diff; at->u_ddiff = m_dd_unused;
    at->v = m_v; at->v_diff = m_v_diff; at->v_ddiff = m_dd_unused;
    at->rho = m_rho;
    return at;
}

/* Checker access to what Set_Slopes_Multi produced: which = 0 values, 1 vertical slopes, 2 slopes of d/d(ax0), 3 slopes of
 * d/d(ax1) of field 0 T / 1 u / 2 v / 3 rho, dense [n0][n1][nz] into out.  Returns the node count (0 on bad arguments). */
int64_t orc_atmo3d_slopes(const orc_atmo* at, int field, int which, double* out) {
    if (!at || !at->grid || field < 0 || field > 3 || which < 0 || which > 3) return 0;
    const ms_grid* g = (const ms_grid*)at->grid;
    const size_t N = (size_t)g->n0 * g->n1 * g->nz;
    const ms_field* F = &g->F[field];
    const double* src = which == 0 ? F->f : (which == 1 ? F->s : (which == 2 ? F->sa : F->sb));
    memcpy(out, src, N * sizeof(double));
    return (int64_t)N;
}

/* Load_G2S_Multi: read `prefix<idx>.met` for every node (Cartesian idx = i0*n1 + i1, :154; Global idx = it*np + ip, :165)
 * plus the two node-coordinate files; winds m/s -> km/s with the ground taper (width 0.05 Cartesian, 0.2 Global; the
 * reference's z_grnd is 0 at load time).  Global lat/lon files are degrees -> radians.  Outputs as orc_atmo3d_create wants. */
int orc_load_met_grid(const char* prefix, const char* loc0, const char* loc1, const char* format, int global,
                      int cap0, int cap1, int capz, int* n0, int* n1, int* nz,
                      double* ax0, double* ax1, double* axz, double* T, double* u, double* v, double* rho) {
    const double Pi = 3.141592653589793238462643;
    int fmt = !strncmp(format, "zTuvdp", 6) ? 0 : (!strncmp(format, "zuvwTdp", 7) ? 1 : -1);
    if (fmt < 0) return GEOAC_ERR_BAD_ARG;
    FILE* f = fopen(loc0, "r"); if (!f) return GEOAC_ERR_IO;
    int c0 = 0; while (c0 < cap0 && fscanf(f, "%lf", &ax0[c0]) == 1) c0++;
    fclose(f);
    f = fopen(loc1, "r"); if (!f) return GEOAC_ERR_IO;
    int c1 = 0; while (c1 < cap1 && fscanf(f, "%lf", &ax1[c1]) == 1) c1++;
    fclose(f);
    if (global) { for (int i = 0; i < c0; i++) ax0[i] *= Pi / 180.0; for (int i = 0; i < c1; i++) ax1[i] *= Pi / 180.0; }
    int cz = -1;
    char path[4096];
    for (int i = 0; i < c0; i++) for (int j = 0; j < c1; j++) {
        snprintf(path, sizeof path, "%s%i.met", prefix, i * c1 + j);
        f = fopen(path, "r"); if (!f) return GEOAC_ERR_IO;
        int k = 0; double zz, tt, uu, vv, rr, t1, t2;
        for (;;) {
            int ok = fmt == 0 ? fscanf(f, "%lf %lf %lf %lf %lf %lf", &zz, &tt, &uu, &vv, &rr, &t1) == 6
                              : fscanf(f, "%lf %lf %lf %lf %lf %lf %lf", &zz, &uu, &vv, &t1, &tt, &rr, &t2) == 7;
            if (!ok || k >= capz || (cz >= 0 && k >= cz)) break;
            double arg;
            if (global) { double r = zz + 6370.0; arg = -(r - 6370.0 - 0.0) / 0.2; }
            else        arg = -(zz - 0.0) / 0.05;
            uu *= (2.0 / (1.0 + exp(arg)) - 1.0) / 1000.0;
            vv *= (2.0 / (1.0 + exp(arg)) - 1.0) / 1000.0;
            size_t id = ((size_t)i * c1 + j) * (size_t)(cz >= 0 ? cz : capz) + k;
            axz[k] = zz; T[id] = tt; u[id] = uu; v[id] = vv; rho[id] = rr;
            k++;
        }
        fclose(f);
        if (cz < 0) {           /* first file fixes nz; compact the first column from stride capz to stride nz (no-op: it is column 0) */
            cz = k;
        } else if (k != cz) return GEOAC_ERR_IO;
    }
    *n0 = c0; *n1 = c1; *nz = cz;
    return (c0 >= 2 && c1 >= 2 && cz >= 3) ? GEOAC_OK : GEOAC_ERR_IO;
}
