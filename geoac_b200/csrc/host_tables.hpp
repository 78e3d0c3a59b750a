// host_tables.hpp -- host-side construction of the range-dependent node tables (plain C++, no CUDA).
// Included by capi.cu (the product) and by tests/host_emul (the g++ debugging build of the device code).
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>
#include "mspline.cuh"

namespace geoac {

// ---- range-dependent tables: Set_Slopes_Multi (G2S_MultiDimSpline3D.cpp:306-425, G2S_GlobalMultiDimSpline3D.cpp:313-431) ----
// One Thomas solve per column for f and for the node finite differences d/d(ax0), d/d(ax1); same recurrences as the
// reference so the slope tables are bit-identical, including the Global file's dfdt[i]-dfdt[i+1] slip.
inline void column_slopes(const double* z, int n, const double* vals, int vstride, double* out, int ostride, bool shifted,
                          std::vector<double>& nc, std::vector<double>& nd) {
    auto v = [&](int i) { return vals[(size_t)i * vstride]; };
    double ai, bi, ci, di;
    bi = 2.0 / (z[1] - z[0]); ci = 1.0 / (z[1] - z[0]);
    di = 3.0 * (v(1) - v(0)) / std::pow(z[1] - z[0], 2);
    nc[0] = ci / bi; nd[0] = di / bi;
    for (int i = 1; i < n - 1; i++) {
        ai = 1.0 / (z[i] - z[i - 1]);
        bi = 2.0 * (1.0 / (z[i] - z[i - 1]) + 1.0 / (z[i + 1] - z[i]));
        ci = 1.0 / (z[i + 1] - z[i]);
        if (shifted) di = 3.0 * ((v(i) - v(i + 1)) / std::pow(z[i] - z[i - 1], 2) + (v(i + 1) - v(i)) / std::pow(z[i + 1] - z[i], 2));
        else         di = 3.0 * ((v(i) - v(i - 1)) / std::pow(z[i] - z[i - 1], 2) + (v(i + 1) - v(i)) / std::pow(z[i + 1] - z[i], 2));
        nc[i] = ci / (bi - nc[i - 1] * ai);
        nd[i] = (di - nd[i - 1] * ai) / (bi - nc[i - 1] * ai);
    }
    ai = 1.0 / (z[n - 1] - z[n - 2]); bi = 2.0 / (z[n - 1] - z[n - 2]);
    di = 3.0 * (v(n - 1) - v(n - 2)) / std::pow(z[n - 1] - z[n - 2], 2);
    nd[n - 1] = (di - nd[n - 2] * ai) / (bi - nc[n - 2] * ai);
    out[(size_t)(n - 1) * ostride] = nd[n - 1];
    for (int i = n - 2; i >= 0; i--) out[(size_t)i * ostride] = nd[i] - nc[i] * out[(size_t)(i + 1) * ostride];
}


// Axis records (mspline.cuh): { x[k], 1/(x[k+1]-x[k]), 1/(x[k+1]-x[max(k-1,0)]), 1/(x[min(k+2,n-1)]-x[k]) } per node index.
inline void build_axis_records(const double* x, int n, std::vector<double>& out) {
    out.assign((size_t)n * AX, 0.0);
    for (int k = 0; k < n; k++) {
        const int km = std::max(k - 1, 0), kp = std::min(k + 2, n - 1), k1 = std::min(k + 1, n - 1);
        out[(size_t)k * AX] = x[k];
        if (k1 > k) {
            out[(size_t)k * AX + 1] = 1.0 / (x[k1] - x[k]);
            out[(size_t)k * AX + 2] = 1.0 / (x[k1] - x[km]);
            out[(size_t)k * AX + 3] = 1.0 / (x[kp] - x[k]);
        }
    }
}

// Interleaved device layout (mspline.cuh): tuv[node][level][field]{f, slope, d/dax0 slope, d/dax1 slope, df/dax0, df/dax1},
// rho[node][level]{f, slope}.
// z receives the vertical coordinate as the kernel sees it (altitude, or r = altitude + r_earth for the Global variant).
inline void build_grid_tables(bool glob, int n0, int n1, int nz, const double* ax0, const double* ax1, const double* axz,
                              const double* T, const double* u, const double* v, const double* rho,
                              std::vector<double>& z, std::vector<double>& tuv, std::vector<double>& rh) {
    const size_t nodes = (size_t)n0 * n1 * nz;
    z.resize(nz);
    for (int k = 0; k < nz; k++) z[k] = axz[k] + (glob ? kREarth : 0.0);                      // r_vals[nr] += r_earth
    tuv.assign(nodes * MS_STRIDE, 0.0); rh.assign(nodes * 2, 0.0);
    std::vector<double> da(nz), db(nz), nc(nz), nd(nz);
    const double* src[4] = { T, u, v, rho };
    for (int F = 0; F < 4; F++) {
        const double* f = src[F];
        for (int i = 0; i < n0; i++) for (int j = 0; j < n1; j++) {
            const size_t col = ((size_t)i * n1 + j) * nz;
            const int iu = std::min(i + 1, n0 - 1), id = std::max(i - 1, 0), ju = std::min(j + 1, n1 - 1), jd = std::max(j - 1, 0);
            double* o = (F < 3) ? &tuv[col * MS_STRIDE + MS_FIELD * F] : &rh[col * 2];
            const int os = (F < 3) ? MS_STRIDE : 2;
            for (int k = 0; k < nz; k++) o[(size_t)k * os] = f[col + k];
            column_slopes(z.data(), nz, f + col, 1, o + 1, os, false, nc, nd);
            if (F < 3) {
                for (int k = 0; k < nz; k++) {
                    da[k] = (f[((size_t)iu * n1 + j) * nz + k] - f[((size_t)id * n1 + j) * nz + k]) / (ax0[iu] - ax0[id]);
                    db[k] = (f[((size_t)i * n1 + ju) * nz + k] - f[((size_t)i * n1 + jd) * nz + k]) / (ax1[ju] - ax1[jd]);
                }
                column_slopes(z.data(), nz, da.data(), 1, o + 2, os, glob, nc, nd);
                column_slopes(z.data(), nz, db.data(), 1, o + 3, os, glob, nc, nd);
                for (int k = 0; k < nz; k++) { o[(size_t)k * os + 4] = da[k]; o[(size_t)k * os + 5] = db[k]; }
            }
        }
    }
}

}  // namespace geoac
