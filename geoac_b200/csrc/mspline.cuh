// mspline.cuh -- range-dependent atmosphere sampler for one thread per ray.
//
// What the reference computes (Code/Atmo/G2S_MultiDimSpline3D.cpp:1156-1593, G2S_GlobalMultiDimSpline3D.cpp:1047-1461;
// behavioural spec in SURVEY App. E): natural cubic splines in the vertical at every horizontal node, and per query five
// bicubic Hermite patches whose corner data are centred finite differences of vertical-spline values, each pushed
// through a dense 16x16 matrix -- 148 vertical-spline calls and ~95 k floating-point operations per field triple.
//
// What this file does instead: the Hermite patch with finite-difference corner data is a TENSOR PRODUCT of 1-D
// operators, so every output is a small bilinear form over the 4x4 node block
//     HX(Q) = h00 Q1 + h01 Q2 + dx (h10 r0 (Q2-Q0) + h11 r1 (Q3-Q1))         (1-D Hermite with FD slopes)
//     FX(V,G) = h00 r0 (V2-V0) + h01 r1 (V3-V1) + dx (h10 r0 (G2-G0) + h11 r1 (G3-G1))   (values dV/dx, slopes dG/dx)
// applied along ax0 inside a row and accumulated over the four rows with the matching ax1 weights.  One pass over the
// 16 nodes yields the value, the gradient and the six second derivatives of T, u and v (~4 k flops instead of ~95 k),
// identical to the reference up to rounding.  The reference's formula slips are kept: the
// Cartesian wrappers / d2f/dz2 block scale the ax1 slope data by dx, the Global vertical-derivative columns drop their
// leading term, and Global second derivatives are not divided by the cell size.
//
// Node data layout in HBM (read through L1/L2; a ray stays in one cell for tens of steps):
//     tuv[((i0*n1 + i1)*nz + k)*18 + 6*field + {0: f, 1: df/dz slope, 2: d(df/dax0)/dz slope, 3: d(df/dax1)/dz slope,
//                                              4: df/dax0, 5: df/dax1 (centred node differences, one-sided at the rim)}]
//     rho[((i0*n1 + i1)*nz + k)*2  + {0: f, 1: slope}]
// The node differences are what the reference's Eval_Vert_Spline_dfdx / dfdy recompute from the neighbouring columns on
// every query (G2S_MultiDimSpline3D.cpp:476-562); storing them with the node makes a column self-contained: three
// aligned 16-byte loads per field and level, no neighbour look-ups, and the two levels a column needs are 2 x 144
// contiguous bytes for all three fields.
#pragma once
#include "core.cuh"

namespace geoac {

struct Grid3D {
    const double* tuv;
    const double* rho;
    const double* ax0; const double* ax1; const double* axz;
    int n0, n1, nz;
    double amin, amax, bmin, bmax, zmin, zmax;
    double* scratch;        // per-thread sampler output block (MS_SCRATCH doubles): shared memory on the device
};
constexpr int MS_SCRATCH = 30;

struct Cur3 { int ka, kb, kz; };

constexpr int MS_STRIDE = 18;      // doubles per node and level in `tuv`
constexpr int MS_FIELD = 6;        // doubles per field inside a node record

// Find_Segment from a cold cursor (both reference files): alternate a search from the bottom and from the top, so a point
// exactly on knot m lands in cell m-1 in the lower half of the axis and in cell m in the upper half.
GEOAC_HD int ms_find_cold(const double* x, int n, double xq) {
    for (int i = 0; i < n; i++) {
        if (xq >= x[i] && xq <= x[i + 1]) return i;
        if (xq >= x[n - 2 - i] && xq < x[n - 1 - i]) return n - 2 - i;
    }
    return 0;
}
// warm cursor: stay in the current cell when the point is on one of its knots (the reference's first test)
GEOAC_HD int ms_find_warm(const double* x, int n, double xq, int k) {
    k = (k < 0) ? 0 : ((k > n - 2) ? n - 2 : k);
    while (xq < x[k]) --k;
    while (xq > x[k + 1]) ++k;
    return k;
}

struct MsAxis {
    unsigned off[4];                     // element offsets of the 4 slot nodes (k-1, k, k+1, k+2 clamped)
    double r0, r1;                       // 1 / (x[up] - x[dn]) of the finite differences centred on nodes k and k+1
    double d, t;                         // cell width, scaled coordinate
};

GEOAC_HD void ms_axis(MsAxis& A, const double* x, int n, int k, double xq, unsigned stride) {
    const int km = (k - 1 < 0) ? 0 : k - 1, kp = (k + 2 > n - 1) ? n - 1 : k + 2;
    A.off[0] = (unsigned)km * stride; A.off[1] = (unsigned)k * stride; A.off[2] = (unsigned)(k + 1) * stride; A.off[3] = (unsigned)kp * stride;
    const double xm = x[km], x0 = x[k], x1 = x[k + 1], xp = x[kp];
    A.r0 = 1.0 / (x1 - xm);              // node k:   up = k+1, dn = max(k-1, 0)
    A.r1 = 1.0 / (xp - x0);              // node k+1: up = min(k+2, n-1), dn = k
    A.d = x1 - x0;
    A.t = (xq - x0) / A.d;
}

// 1-D weights of one axis: Hermite basis at t (and its t-derivative) combined with the corner finite differences
struct MsW {
    double h00, h01, S0, S1;       // value basis; slope basis times the cell width
    double p, q, P, Q;             // h00 r0, h01 r1, S0 r0, S1 r1
    double e00, e01, T0, T1;       // derivatives of the above with respect to t
    double pd, qd, Pd, Qd;
};
GEOAC_HD void ms_weights(MsW& w, const MsAxis& A, double slope_scale) {
    const double t = A.t, u = 1.0 - t;
    w.h00 = (1.0 + 2.0 * t) * u * u; w.h01 = t * t * (3.0 - 2.0 * t);
    w.S0 = slope_scale * (t * u * u); w.S1 = slope_scale * (t * t * (t - 1.0));
    w.e00 = 6.0 * t * (t - 1.0); w.e01 = -w.e00;
    w.T0 = slope_scale * (u * (1.0 - 3.0 * t)); w.T1 = slope_scale * (t * (3.0 * t - 2.0));
    const double r0 = A.r0, r1 = A.r1;
    w.p = w.h00 * r0; w.q = w.h01 * r1; w.P = w.S0 * r0; w.Q = w.S1 * r1;
    w.pd = w.e00 * r0; w.qd = w.e01 * r1; w.Pd = w.T0 * r0; w.Qd = w.T1 * r1;
}
// row operators along one axis (Q0..Q3 = quantity on the four slot nodes)
GEOAC_HD double ms_HX(const MsW& w, double Q0, double Q1, double Q2, double Q3)  { return w.h00 * Q1 + w.h01 * Q2 + w.P * (Q2 - Q0) + w.Q * (Q3 - Q1); }
GEOAC_HD double ms_HXd(const MsW& w, double Q0, double Q1, double Q2, double Q3) { return w.e00 * Q1 + w.e01 * Q2 + w.Pd * (Q2 - Q0) + w.Qd * (Q3 - Q1); }

// vertical position shared by every column of a query.  The vertical Hermite spline is LINEAR in a column's four data
// (f_k, f_k+1, s_k, s_k+1), so its value / first / second derivative at this position are dot products with coefficient
// triples that depend on the position only -- computed once per query, 3 FMAs per column and quantity instead of ~9:
//     V   = f_k + cV1 (f_k+1 - f_k) + cVa s_k + cVb s_k+1          (same polynomial as G2S_MultiDimSpline3D.cpp:476-562)
//     V'  =       cD1 (f_k+1 - f_k) + cDa s_k + cDb s_k+1
//     V'' =       cE1 (f_k+1 - f_k) + cEa s_k + cEb s_k+1
struct MsZ { int kz; double cV1, cVa, cVb, cD1, cDa, cDb, cE1, cEa, cEb, cG1; };
template <bool GLOBAL>
GEOAC_HD void ms_zpos(MsZ& Z, const Grid3D& g, double z, int kz) {
    const double z0 = g.axz[kz];
    const double h = g.axz[kz + 1] - z0, invh = 1.0 / h;
    const double X = (z - z0) * invh, omX = 1.0 - X, XomX = X * omX, om2X = 1.0 - 2.0 * X;
    Z.kz = kz;
    Z.cV1 = X - XomX * om2X;            Z.cVa = XomX * omX * h;           Z.cVb = -(XomX * X * h);
    Z.cD1 = (1.0 - om2X * om2X + 2.0 * XomX) * invh;  Z.cDa = om2X * omX - XomX;  Z.cDb = -(om2X * X + XomX);
    const double inv2 = 2.0 * invh * invh;
    Z.cE1 = 3.0 * om2X * inv2;          Z.cEa = (3.0 * X - 2.0) * h * inv2;  Z.cEb = (3.0 * X - 1.0) * h * inv2;
    // vertical derivative of a finite-difference column: the Global file drops the leading (d1 - d0)/h term (App. A-9)
    Z.cG1 = GLOBAL ? Z.cD1 - invh : Z.cD1;
}
GEOAC_HD double ms_v(const MsZ& Z, double f0, double df, double s0, double s1)  { return fma(Z.cV1, df, fma(Z.cVa, s0, fma(Z.cVb, s1, f0))); }
GEOAC_HD double ms_vz(const MsZ& Z, double df, double s0, double s1)            { return fma(Z.cD1, df, fma(Z.cDa, s0, Z.cDb * s1)); }
GEOAC_HD double ms_vzz(const MsZ& Z, double df, double s0, double s1)           { return fma(Z.cE1, df, fma(Z.cEa, s0, Z.cEb * s1)); }
GEOAC_HD double ms_gz(const MsZ& Z, double dd, double s0, double s1)            { return fma(Z.cG1, dd, fma(Z.cDa, s0, Z.cDb * s1)); }

// the six data of one node and level for one field (48 aligned bytes): f, df/dz slope, the vertical slopes of the two
// node differences, and the node differences df/dax0, df/dax1 themselves
struct Node6 { double f, s, sa, sb, da, db; };
GEOAC_HD Node6 ms_ld6(const double* p) {
    const Pair a = ld_pair(p), b = ld_pair(p + 2), c = ld_pair(p + 4);
    Node6 n; n.f = a.a; n.s = a.b; n.sa = b.a; n.sb = b.b; n.da = c.a; n.db = c.b; return n;
}

// ---------------------------------------------------------------------------------------------------------------
// Eval_Spline_AllOrder1 / AllOrder2 for T, u and v in one pass.  out[field][..] in GRID-axis order:
//   0 f, 1 d/da, 2 d/db, 3 d/dz, 4 d2/da2, 5 d2/db2, 6 d2/dz2, 7 d2/dadb, 8 d2/dadz, 9 d2/dbdz   (a = ax0, b = ax1, z = vertical)
// ---------------------------------------------------------------------------------------------------------------
template <bool GLOBAL, bool ORDER2>
GEOAC_HD void ms_sample_tuv(const Grid3D& g, double a_in, double b_in, double z_in, Cur3& cur, double (&out)[3][10]) {
    const double a = clampd(a_in, g.amin, g.amax), b = clampd(b_in, g.bmin, g.bmax), z = clampd(z_in, g.zmin, g.zmax);
    cur.ka = ms_find_warm(g.ax0, g.n0, a, cur.ka);
    cur.kb = ms_find_warm(g.ax1, g.n1, b, cur.kb);
    cur.kz = ms_find_warm(g.axz, g.nz, z, cur.kz);
    MsAxis A, B; MsZ Z;
    ms_axis(A, g.ax0, g.n0, cur.ka, a, (unsigned)(g.n1 * g.nz * MS_STRIDE));
    ms_axis(B, g.ax1, g.n1, cur.kb, b, (unsigned)(g.nz * MS_STRIDE));
    ms_zpos<GLOBAL>(Z, g, z, cur.kz);
    MsW wa, wb;
    ms_weights(wa, A, A.d);
    ms_weights(wb, B, B.d);
    // d2f/dz2 block: the Cartesian file scales the ax1 slope data by dx (App. A-8); Global uses dp
    const double qs = GLOBAL ? 1.0 : A.d / B.d;
    const unsigned kofs = (unsigned)cur.kz * MS_STRIDE;

#pragma unroll 1
    for (int F = 0; F < 3; F++) {
        double acc[10];
#pragma unroll
        for (int i = 0; i < 10; i++) acc[i] = 0.0;
        const double* base = g.tuv + kofs + MS_FIELD * F;
#pragma unroll
        for (int jb = 0; jb < 4; jb++) {
            double V[4], Vz[4], Vzz[4], Ga[4], Gb[4], Gaz[4], Gbz[4];
#pragma unroll
            for (int ia = 0; ia < 4; ia++) {
                const double* n0p = base + (A.off[ia] + B.off[jb]);
                const Node6 lo = ms_ld6(n0p), hi = ms_ld6(n0p + MS_STRIDE);
                const double df = hi.f - lo.f, dda = hi.da - lo.da, ddb = hi.db - lo.db;
                V[ia] = ms_v(Z, lo.f, df, lo.s, hi.s);
                Vz[ia] = ms_vz(Z, df, lo.s, hi.s);
                if (ORDER2) Vzz[ia] = ms_vzz(Z, df, lo.s, hi.s);
                Ga[ia] = ms_v(Z, lo.da, dda, lo.sa, hi.sa);          // column of the ax0 node difference (Eval_Vert_Spline_dfdx / ddfdxdz)
                Gaz[ia] = ms_gz(Z, dda, lo.sa, hi.sa);
                Gb[ia] = ms_v(Z, lo.db, ddb, lo.sb, hi.sb);          // column of the ax1 node difference
                Gbz[ia] = ms_gz(Z, ddb, lo.sb, hi.sb);
            }
            // ax1 weights of this row (slot jb): tensor weight, its derivative, FD-of-values weight, FD-of-slopes weight
            const double ey  = (jb == 1) ? wb.h00 : ((jb == 2) ? wb.h01 : 0.0);
            const double eyd = (jb == 1) ? wb.e00 : ((jb == 2) ? wb.e01 : 0.0);
            const double sy  = (jb == 1) ? wb.S0 : ((jb == 2) ? wb.S1 : 0.0);
            const double syd = (jb == 1) ? wb.T0 : ((jb == 2) ? wb.T1 : 0.0);
            const double al  = (jb == 0) ? -wb.p  : ((jb == 1) ? -wb.q  : ((jb == 2) ? wb.p  : wb.q));
            const double ald = (jb == 0) ? -wb.pd : ((jb == 1) ? -wb.qd : ((jb == 2) ? wb.pd : wb.qd));
            const double be  = (jb == 0) ? -wb.P  : ((jb == 1) ? -wb.Q  : ((jb == 2) ? wb.P  : wb.Q));
            const double bed = (jb == 0) ? -wb.Pd : ((jb == 1) ? -wb.Qd : ((jb == 2) ? wb.Pd : wb.Qd));
            const double wy = ey + be, wyd = eyd + bed;

            const double Fv = ms_HX(wa, V[0], V[1], V[2], V[3]);
            const double dV0 = V[2] - V[0], dV1 = V[3] - V[1], dG0 = Ga[2] - Ga[0], dG1 = Ga[3] - Ga[1];
            const double FXv = wa.p * dV0 + wa.q * dV1 + wa.P * dG0 + wa.Q * dG1;
            const double FGb = ms_HX(wa, Gb[0], Gb[1], Gb[2], Gb[3]);
            const double EVz = wa.h00 * Vz[1] + wa.h01 * Vz[2];
            const double BVz = wa.P * (Vz[2] - Vz[0]) + wa.Q * (Vz[3] - Vz[1]);
            const double GXZ = wa.S0 * Gaz[1] + wa.S1 * Gaz[2];          // used only on the corner rows (ey != 0)
            const double GYZ = wa.h00 * Gbz[1] + wa.h01 * Gbz[2];
            acc[0] += wy * Fv;
            acc[1] += wy * FXv;
            acc[2] += al * Fv + be * FGb;
            acc[3] += ey * (EVz + GXZ) + be * BVz + sy * GYZ;
            if (ORDER2) {
                const double FXd = wa.pd * dV0 + wa.qd * dV1 + wa.Pd * dG0 + wa.Qd * dG1;
                const double EVzd = wa.e00 * Vz[1] + wa.e01 * Vz[2];
                const double BVzd = wa.Pd * (Vz[2] - Vz[0]) + wa.Qd * (Vz[3] - Vz[1]);
                const double GXZd = wa.T0 * Gaz[1] + wa.T1 * Gaz[2];
                const double GYZd = wa.e00 * Gbz[1] + wa.e01 * Gbz[2];
                acc[4] += wy * FXd;
                acc[5] += ald * Fv + bed * FGb;
                {   // d2f/dz2 block: only its Py data carry the dx-for-dy slip, the Pxy data are scaled by dx*dy
                    const double EVzz = wa.h00 * Vzz[1] + wa.h01 * Vzz[2];
                    const double BVzz = wa.P * (Vzz[2] - Vzz[0]) + wa.Q * (Vzz[3] - Vzz[1]);
                    acc[6] += ey * (EVzz + BVzz) + (be * qs) * EVzz + be * BVzz;
                }
                acc[7] += wyd * FXv;
                acc[8] += ey * (EVzd + GXZd) + be * BVzd + sy * GYZd;
                acc[9] += eyd * (EVz + GXZ) + bed * BVz + syd * GYZ;
            }
        }
        if (ORDER2 && !GLOBAL) {            // Global leaves the second derivatives in scaled units (App. A-9)
            const double ida = 1.0 / A.d, idb = 1.0 / B.d;
            acc[4] *= ida; acc[5] *= idb; acc[7] *= idb; acc[8] *= ida; acc[9] *= idb;
        }
#pragma unroll
        for (int i = 0; i < 10; i++) out[F][i] = acc[i];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The scalar wrappers c(), u(), v(), rho() (Eval_Spline_f: Cartesian scales the ax1 slope data by dx, App. A-8) and the
// vertical derivative wrappers c_diff / u_diff / v_diff (Eval_Spline_df with the same scaling) -- used by travel time,
// absorption, amplitude, reflection and the per-launch invariants.  vals: T, u, v, rho;  dz: dT/dz, du/dz, dv/dz.
// ---------------------------------------------------------------------------------------------------------------
template <bool GLOBAL, bool WITH_RHO, bool WITH_DZ>
GEOAC_HD void ms_wrappers(const Grid3D& g, double a_in, double b_in, double z_in, Cur3& cur, double (&vals)[4], double (&dz)[3]) {
    const double a = clampd(a_in, g.amin, g.amax), b = clampd(b_in, g.bmin, g.bmax), z = clampd(z_in, g.zmin, g.zmax);
    cur.ka = ms_find_warm(g.ax0, g.n0, a, cur.ka);
    cur.kb = ms_find_warm(g.ax1, g.n1, b, cur.kb);
    cur.kz = ms_find_warm(g.axz, g.nz, z, cur.kz);
    MsAxis A, B; MsZ Z;
    ms_axis(A, g.ax0, g.n0, cur.ka, a, (unsigned)(g.n1 * g.nz * MS_STRIDE));
    ms_axis(B, g.ax1, g.n1, cur.kb, b, (unsigned)(g.nz * MS_STRIDE));
    ms_zpos<GLOBAL>(Z, g, z, cur.kz);
    MsW wa, wb;
    ms_weights(wa, A, A.d);
    ms_weights(wb, B, GLOBAL ? B.d : A.d);              // the quirk: ax1 slope data times dx
    const unsigned kofs = (unsigned)cur.kz * MS_STRIDE;
#pragma unroll 1
    for (int F = 0; F < (WITH_RHO ? 4 : 3); F++) {
        const bool is_rho = (F == 3);
        const double* base = is_rho ? g.rho : g.tuv + MS_FIELD * F;
        const int lvl = is_rho ? 2 : MS_STRIDE;          // doubles between vertical levels
        const unsigned ko = is_rho ? (unsigned)cur.kz * 2u : kofs;
        const unsigned shr = is_rho ? (unsigned)(MS_STRIDE / 2) : 1u;   // node offsets were built for the tuv layout: /9 for the 2-double one
        double accv = 0.0, accz = 0.0;
#pragma unroll
        for (int jb = 0; jb < 4; jb++) {
            double V[4], Vz[4], Gaz[4], Gbz[4];
#pragma unroll
            for (int ia = 0; ia < 4; ia++) {
                const unsigned o = (A.off[ia] + B.off[jb]) / shr;
                const double* n0p = base + ko + o;
                const Pair lo = ld_pair(n0p), hi = ld_pair(n0p + lvl);            // (f, slope) at both levels
                const double df = hi.a - lo.a;
                V[ia] = ms_v(Z, lo.a, df, lo.b, hi.b);
                if (WITH_DZ && !is_rho) {
                    Vz[ia] = ms_vz(Z, df, lo.b, hi.b);
                    const Pair slo = ld_pair(n0p + 2), shi = ld_pair(n0p + lvl + 2);  // (sa, sb) at both levels
                    const Pair dlo = ld_pair(n0p + 4), dhi = ld_pair(n0p + lvl + 4);  // (da, db) at both levels
                    Gaz[ia] = ms_gz(Z, dhi.a - dlo.a, slo.a, shi.a);
                    Gbz[ia] = ms_gz(Z, dhi.b - dlo.b, slo.b, shi.b);
                }
            }
            const double ey = (jb == 1) ? wb.h00 : ((jb == 2) ? wb.h01 : 0.0);
            const double sy = (jb == 1) ? wb.S0 : ((jb == 2) ? wb.S1 : 0.0);
            const double be = (jb == 0) ? -wb.P : ((jb == 1) ? -wb.Q : ((jb == 2) ? wb.P : wb.Q));
            // the Pxy block of Eval_Spline_f / _df is scaled by dx*dy, only the Py block by the slipped scale: separate the two
            const double be_true = GLOBAL ? be : be * (B.d / A.d);
            {
                const double EV = wa.h00 * V[1] + wa.h01 * V[2];
                const double BV = wa.P * (V[2] - V[0]) + wa.Q * (V[3] - V[1]);
                accv += ey * (EV + BV) + be * EV + be_true * BV;
            }
            if (WITH_DZ && !is_rho) {
                const double EVz = wa.h00 * Vz[1] + wa.h01 * Vz[2];
                const double BVz = wa.P * (Vz[2] - Vz[0]) + wa.Q * (Vz[3] - Vz[1]);
                const double GXZ = wa.S0 * Gaz[1] + wa.S1 * Gaz[2];
                const double GYZ = wa.h00 * Gbz[1] + wa.h01 * Gbz[2];
                accz += ey * (EVz + GXZ) + be_true * BVz + sy * GYZ;
            }
        }
        vals[F] = accv;
        if (WITH_DZ && !is_rho) dz[F] = accz;
    }
}

}  // namespace geoac
