// core.cuh -- shared device-side building blocks of the batched ray tracer.
//
// Everything here is written from scratch for a one-thread-per-ray FP64 kernel: atmosphere samples are
// evaluated ONCE per RK4 stage (the reference re-evaluates c(), u(), ... dozens of times per stage), divisions are
// turned into a handful of reciprocals per stage, per-launch invariants live in a constant block.  The math follows
//   splines      : reference Code/Atmo/G2S_Spline1D.cpp:245-281 (Hermite "slopes" form of the natural cubic spline)
//   c, c', c''   : Code/Atmo/G2S_Spline1D.cpp:334-358       (c = sqrt(gamR*T), chain rule)
//   absorption   : Code/Atmo/Atmo_State.Absorption.cpp:14-143 (Sutherland & Bass 2004)
// Functions are GEOAC_HD so that tests/host_emul can compile the same per-ray logic with g++ as a debugging aid;
// the product library only ever instantiates them inside __global__ kernels.
#pragma once
#include <math.h>
#include <stdint.h>
#include "../../include/geoac_b200.h"

#ifdef __CUDACC__
#define GEOAC_HD __host__ __device__ __forceinline__
#define GEOAC_HD_NOINLINE __host__ __device__ __noinline__
#else
#define GEOAC_HD inline
#define GEOAC_HD_NOINLINE
#endif

namespace geoac {

constexpr double kPi   = 3.141592653589793238462643;   // Code/GeoAc/GeoAc.Parameters.cpp:28-31
constexpr double kGam  = 1.4;
constexpr double kR    = 287.05;
constexpr double kGamR = 0.00040187;                    // Code/Atmo/G2S_Spline1D.cpp:332
constexpr double kREarth = 6370.0;                      // Code/Atmo/G2S_GlobalSpline1D.cpp:35

#if defined(__CUDA_ARCH__)
GEOAC_HD double g_rsqrt(double x) { return rsqrt(x); }
GEOAC_HD double g_rcbrt(double x) { return rcbrt(x); }
GEOAC_HD double g_exp10(double x) { return exp10(x); }
GEOAC_HD double g_rcp(double x)   { return 1.0 / x; }
#else
GEOAC_HD double g_rsqrt(double x) { return 1.0 / sqrt(x); }
GEOAC_HD double g_rcbrt(double x) { return 1.0 / cbrt(x); }
GEOAC_HD double g_exp10(double x) { return pow(10.0, x); }
GEOAC_HD double g_rcp(double x)   { return 1.0 / x; }
#endif

// ---------------------------------------------------------------------------------------------------------------
// 1-D table: structure-of-arrays, `n` levels per array, in this order (each array padded to n_pad doubles):
//   x, invh, T, sT, u, su, v, sv, rho, srho          (invh[k] = 1/(x[k+1]-x[k]), invh[n-1] = 0)
// SoA keeps neighbouring levels in neighbouring shared-memory banks; lanes at the same level broadcast.
// ---------------------------------------------------------------------------------------------------------------
enum { TAB_X = 0, TAB_INVH, TAB_T, TAB_ST, TAB_U, TAB_SU, TAB_V, TAB_SV, TAB_RHO, TAB_SRHO, TAB_NARR };

struct Table1D {
    const double* base;     // shared or global memory
    int n, n_pad;
    double xmin, xmax;
    GEOAC_HD const double* arr(int a) const { return base + (size_t)a * n_pad; }
};

struct SegPos { int k; double X, h, invh; };

// locate the interval containing xc (already clamped into [xmin,xmax]) starting from cursor k.
// Same tie-breaking as the reference's Find_Segment (G2S_Spline1D.cpp:202-243): a point on a knot stays in the
// interval the cursor is already in.
GEOAC_HD SegPos seg_locate(const Table1D& t, double xc, int& k) {
    const double* x = t.arr(TAB_X);
    double x0 = x[k], x1 = x[k + 1];
    while (xc < x0) { --k; x1 = x0; x0 = x[k]; }
    while (xc > x1) { ++k; x0 = x1; x1 = x[k + 1]; }
    SegPos s; s.k = k; s.h = x1 - x0; s.invh = t.arr(TAB_INVH)[k]; s.X = (xc - x0) * s.invh;
    return s;
}

GEOAC_HD double clampd(double v, double lo, double hi) { return fmax(fmin(v, hi), lo); }

// value only
GEOAC_HD double spl_f(const double* F, const double* S, const SegPos& p) {
    const double f0 = F[p.k], f1 = F[p.k + 1];
    const double df = f1 - f0;
    const double A = S[p.k] * p.h - df, B = -S[p.k + 1] * p.h + df;
    const double omX = 1.0 - p.X;
    return omX * f0 + p.X * f1 + p.X * omX * (A * omX + B * p.X);
}
// value + first derivative
GEOAC_HD void spl_f1(const double* F, const double* S, const SegPos& p, double& f, double& d1) {
    const double f0 = F[p.k], f1 = F[p.k + 1];
    const double df = f1 - f0;
    const double A = S[p.k] * p.h - df, B = -S[p.k + 1] * p.h + df;
    const double omX = 1.0 - p.X, P = A * omX + B * p.X, XomX = p.X * omX;
    f  = omX * f0 + p.X * f1 + XomX * P;
    d1 = (df + (1.0 - 2.0 * p.X) * P + XomX * (B - A)) * p.invh;
}
// value + first + second derivative
GEOAC_HD void spl_f2(const double* F, const double* S, const SegPos& p, double& f, double& d1, double& d2) {
    const double f0 = F[p.k], f1 = F[p.k + 1];
    const double df = f1 - f0;
    const double A = S[p.k] * p.h - df, B = -S[p.k + 1] * p.h + df;
    const double omX = 1.0 - p.X, P = A * omX + B * p.X, XomX = p.X * omX;
    f  = omX * f0 + p.X * f1 + XomX * P;
    d1 = (df + (1.0 - 2.0 * p.X) * P + XomX * (B - A)) * p.invh;
    d2 = 2.0 * (B - 2.0 * A + (A - B) * 3.0 * p.X) * (p.invh * p.invh);
}

// thermodynamic sound speed and its vertical derivatives from T, T', T''
struct SoundSpeed { double c, inv_c, dc, ddc; };
GEOAC_HD SoundSpeed sound_speed2(double T, double dT, double ddT) {
    SoundSpeed s;
    const double gT = kGamR * T;
    s.inv_c = g_rsqrt(gT);
    s.c = gT * s.inv_c;
    const double hg = 0.5 * kGamR * s.inv_c;                  // gamR/(2c)
    s.dc  = hg * dT;
    s.ddc = hg * (ddT - hg * s.inv_c * dT * dT);              // gamR/(2c) T'' - gamR^2/(4c^3) T'^2
    return s;
}
GEOAC_HD double sound_speed0(double T) { return sqrt(kGamR * T); }

// ---------------------------------------------------------------------------------------------------------------
// per-launch invariants (filled by a one-thread setup kernel from the same device routines, then read-only)
// ---------------------------------------------------------------------------------------------------------------
// reference state of the Sutherland-Bass model (T_o, P_o): per launch for the Cartesian variants and the stratified
// Global one, per step for Global.RngDep (its reference point follows the ray's latitude / longitude, App. A-14)
struct SBRef { double invTo, cbrtTo, visc_num, invPo; };

struct LaunchConsts {
    // parameters (copy of geoac_params + derived)
    double ds_min, ds_max, vert_limit, range_limit, z_grnd, tweak_abs, freq;
    double box_min[2], box_max[2];
    double src[3];
    int32_t bounces, calc_amp, seg_mode, step_limit, per_bounce_zmax, pad0;
    double ground;            // z_grnd (Cartesian) or r_earth + z_grnd (Global)
    // source / ground state
    double c_src, u_src, v_src, rho_src;      // at the source point
    double c_000;                             // c(0,0,0)  (3D travel time, App. A-4)
    double c_gnd, rho_gnd;                    // at z_grnd (2D amplitude / reflection)
    double dc_gnd, du_gnd, dv_gnd;            // vertical derivatives at z_grnd (stratified reflection)
    // Sutherland-Bass invariants
    SBRef sb;                                 // 1/T_o, T_o^(1/3), (1 + S/T_o), 1/P_o
    double sb_w;                              // 2*pi*freq
};

// Sutherland-Bass absorption [dB/km] at altitude z [km] with local sound speed c [km/s] (and 1/c) and density rho.
// Restructured from Atmo_State.Absorption.cpp:14-143: constants folded, pow(10,.) -> exp10, pow(T,-1/3) -> rcbrt,
// exp(9.17 Tr) = 1/exp(-9.17 Tr), common factors hoisted.  Branch thresholds are the reference's (strict >).
GEOAC_HD_NOINLINE double suthbass_alpha(const LaunchConsts& L, const SBRef& R, double z, double c, double inv_c, double rho) {
    const double mu_o = 18.192E-6, S = 117.0;
    const double c1000 = c * 1000.0;
    const double c2 = c1000 * c1000;
    const double T_z = c2 * (1.0 / (kR * kGam));
    const double P_z = rho * c2 * (1000.0 / kGam);
    const double inv_Tz = 1.0 / T_z;
    const double inv_Pz = 1.0 / P_z;

    const double mu_ratio = sqrt(T_z * R.invTo) * (R.visc_num / (1.0 + S * inv_Tz));     // mu/mu_o
    const double mu = mu_o * mu_ratio;
    const double nu = (8.0 * kPi * L.freq * mu) * inv_Pz * (1.0 / 3.0);

    const double z2 = z * z, z3 = z2 * z, z4 = z2 * z2, z5 = z4 * z;
    double X0, X1, X3, X4, X5, X6;
    const double X2 = 0.00040003685104612505;                              // 10^-3.3979
    if (z > 90.) X0 = g_exp10(49.296 - (1.5524 * z) + (1.8714E-2 * z2) - (1.1069E-4 * z3) + (3.199E-7 * z4) - (3.6211E-10 * z5));
    else         X0 = 0.20947393930547747;                                 // 10^-0.67887
    if (z > 76.) X1 = g_exp10((1.3972E-1) - (5.6269E-3 * z) + (3.9407E-5 * z2) - (1.0737E-7 * z3));
    else         X1 = 0.780836309209143;                                   // 10^-0.10744
    if (z > 80.) X3 = g_exp10(-4.234 - (3.0975E-2 * z));
    else         X3 = g_exp10(-19.027 + (1.3093 * z) - (4.6496E-2 * z2) + (7.8543E-4 * z3) - (6.5169E-6 * z4) + (2.1343E-8 * z5));
    if (z > 95.) X4 = g_exp10(-3.2456 + (4.6642E-2 * z) - (2.6894E-4 * z2) + (5.264E-7 * z3));
    else         X4 = g_exp10(-11.195 + (1.5408E-1 * z) - (1.4348E-3 * z2) + (1.0166E-5 * z3));
    X5 = g_exp10(-53.746 + (1.5439 * z) - (1.8824E-2 * z2) + (1.1587E-4 * z3) - (3.5399E-7 * z4) + (4.2609E-10 * z5));
    if (z > 30.) X6 = g_exp10(-4.2563 + (7.6245E-2 * z) - (2.1824E-3 * z2) - (2.3010E-6 * z3) + (2.4265E-7 * z4) - (1.2500E-09 * z5));
    else         X6 = g_exp10(-1.7491 + (4.4986E-2 * z) - (6.8549E-2 * z2) + (5.4639E-3 * z3) - (1.5539E-4 * z4) + (1.5063E-06 * z5));
    const double X_ON = (X0 + X1) * (1.0 / 0.9903);

    const double cb = g_rcbrt(T_z);                                        // T_z^(-1/3)
    const double Zr0 = 54.1 * exp(-17.3 * cb), Zr1 = 63.3 * exp(-16.7 * cb);
    const double Z_rot_ = (Zr0 * Zr1) / (X1 * Zr0 + X0 * Zr1);             // 1/(X1/Zr1 + X0/Zr0)

    const double sigma = 1.091089451179962;                                // 5/sqrt(21)
    const double nn = 0.5237229365663817 * Z_rot_;                         // (4/5) sqrt(3/7) Z_rot_
    const double chi = 0.75 * nn * nu;
    const double cchi = 2.36 * chi;

    const double nu2p1 = 1.0 + nu * nu;
    const double s1 = sqrt(nu2p1);
    const double cchi2p1 = 1.0 + cchi * cchi;
    const double sc = sigma * cchi;
    const double w_c = L.sb_w * inv_c;                                     // 2 pi f / c
    // NB: sqrt(1+nu^2) - 1 cancels catastrophically for small nu (it is 0 or a few ulp below ~60 km at 0.1 Hz).
    // That quantisation IS the reference's observable behaviour (Absorption.cpp:108), so it is reproduced literally.
    const double s1m1 = s1 - 1.0;
    const double a_cl  = w_c * sqrt(0.5 * s1m1 * cchi2p1 / (nu2p1 * (1.0 + sc * sc)));
    const double a_rot = w_c * X_ON * ((sigma * sigma - 1.0) * chi * (0.5 / sigma)) * sqrt(0.5 * (s1 + 1.0) / (nu2p1 * cchi2p1));
    const double a_diff = 0.003 * a_cl;

    const double Tr = cb * R.cbrtTo - 1.0;                              // (T_z/T_o)^(-1/3) - 1
    const double A1 = (X0 + X1) * 24.0 * exp(-9.16 * Tr);
    const double A2 = (X4 + X5) * 2400.0;
    const double B  = 40400.0 * exp(10.0 * Tr);
    const double C  = 0.02 * exp(-11.2 * Tr);
    const double D  = 0.391 * exp(8.41 * Tr);
    const double E  = 9.0 * exp(-19.9 * Tr);
    const double F  = 60000.0;
    const double G  = 28000.0 * exp(-4.17 * Tr);
    const double H  = 22000.0 * exp(-7.68 * Tr);
    const double I  = 15100.0 * exp(-10.4 * Tr);
    const double eJ = exp(-9.17 * Tr);
    const double J  = 11500.0 * eJ;
    const double K  = (8.48E08) / eJ;
    const double Lx = exp(-7.72 * Tr);
    const double ZZ = H * X2 + I * (X0 + 0.5 * X4) + J * (X1 + 0.5 * X5) + K * (X6 + X3);
    const double hu = 100.0 * (X3 + X6);
    const double pm = (P_z * R.invPo) / mu_ratio;                       // (P_z/P_o)(mu_o/mu)
    double fv[4];
    fv[0] = pm * (A1 + A2 + B * hu * (C + hu) * (D + hu));
    fv[1] = pm * (E + F * X3 + G * X6);
    fv[2] = pm * ZZ;
    fv[3] = pm * (1.2E5) * Lx;

    const double th[4] = { 2239.1, 3352.0, 915.0, 1037.0 };
    const double CpR[4] = { 3.5, 3.5, 4.0, 4.0 }, CvR[4] = { 2.5, 2.5, 3.0, 3.0 };
    const double Xm[4] = { X0, X1, X2, X3 };
    const double f2 = L.freq * L.freq;
    double a_vib = 0.0;
#pragma unroll
    for (int m = 0; m < 4; m++) {
        const double r = th[m] * inv_Tz;
        const double e = exp(-r);
        const double ome = 1.0 - e;
        const double C_R = (r * r * e) / (ome * ome);
        const double A_max = (Xm[m] * (kPi / 2) * C_R) / (CpR[m] * (CvR[m] + C_R));
        // (2 f^2/fv)/(1 + (f/fv)^2) = 2 f^2 fv/(fv^2 + f^2)
        a_vib += (A_max * inv_c) * (2.0 * f2 * fv[m] / (fv[m] * fv[m] + f2));
    }
    return (a_cl + a_rot + a_diff + a_vib) * L.tweak_abs * 8.685889;
}

// fill the Sutherland-Bass invariants from the reference state (c, rho at the reference level)
GEOAC_HD void suthbass_ref(SBRef& R, double c_ref, double rho_ref) {
    const double c1000 = c_ref * 1000.0;
    const double T_o = c1000 * c1000 / (kR * kGam);
    const double P_o = rho_ref * (c1000 * c1000) / kGam * 1000.0;
    R.invTo = 1.0 / T_o; R.cbrtTo = cbrt(T_o); R.visc_num = 1.0 + 117.0 / T_o;
    R.invPo = 1.0 / P_o;
}
GEOAC_HD void suthbass_setup(LaunchConsts& L, double c_ref, double rho_ref) {
    suthbass_ref(L.sb, c_ref, rho_ref);
    L.sb_w = 2.0 * kPi * L.freq;
}

}  // namespace geoac
