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
#include <string.h>
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

// ---------------------------------------------------------------------------------------------------------------
// Branch-free FP64 primitives for the per-step hot loop.  libdevice's exp / division / sqrt carry a slow-path test
// (BSSY/BSYNC + branch) per call and materialise every polynomial coefficient with two moves; the hot loop does ~20
// exponentials, ~10 reciprocals and ~8 square roots per RK4 step on operands that are known to be normal, so these
// versions drop the special-case handling and read their coefficients as constant-bank operands of the DFMAs.
// All are accurate to <= 1 ulp-ish (2e-16 relative), far inside the 1e-9 parity budget; the host build (tests/host_emul)
// uses the same polynomial code with exact 1/x and sqrt in place of the MUFU seeds.
// Coefficients: scripts/gen_exp_coeffs.py (degree-11 Chebyshev-node interpolants, max rel err 1.6e-16 / 1.9e-16).
// ---------------------------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
#define GEOAC_CONST_TABLE static __constant__
#else
#define GEOAC_CONST_TABLE static const
#endif
GEOAC_CONST_TABLE double kExpE[12] = { 1.0, 1.0, 0.5000000000000019, 0.1666666666666668, 0.0416666666664881, 0.008333333333319601,
    0.0013888888952314775, 0.00019841269890047113, 2.4801485482328494e-05, 2.755724091857897e-06, 2.763263963904103e-07, 2.5110037605963777e-08 };
GEOAC_CONST_TABLE double kExp10[12] = { 1.0, 2.302585092994046, 2.6509490552392085, 2.034678592293478, 1.1712551489072474, 0.5393829291946926,
    0.20699584964214854, 0.06808936524182622, 0.019597614171033596, 0.005013914586462978, 0.0011576552892216332, 0.00024222554338191172 };

#if defined(__CUDA_ARCH__)
GEOAC_HD double g_scale2(double p, int n) { return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p)); }
GEOAC_HD int    g_lo32(double t) { return __double2loint(t); }
// 1/x for normal x: MUFU seed (>= 20 bits) + one third-order Newton step
GEOAC_HD double g_rcp(double x) {
    double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    return fma(y, fma(e, e, e), y);                       // y (1 + e + e^2): relative error e^3 <= 2^-60 before rounding
}
// 1/sqrt(x) for normal positive x: MUFU seed (~2^-21) + one third-order Newton step
GEOAC_HD double g_rsqrt(double x) {
    double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-(x * y), y, 1.0);
    return fma(y * e, fma(0.375, e, 0.5), y);             // y (1 + e/2 + 3 e^2/8): relative error ~e^3 <= 2^-63 before rounding
}
// sqrt(x) and 1/sqrt(x) together (x > 0 normal)
GEOAC_HD double g_sqrt_rs(double x, double& rs) {
    rs = g_rsqrt(x);
    const double s = x * rs;
    return fma(fma(-s, s, x), 0.5 * rs, s);
}
GEOAC_HD double g_rcbrt(double x) { return rcbrt(x); }
#else
GEOAC_HD double g_scale2(double p, int n) { return ldexp(p, n); }
GEOAC_HD int    g_lo32(double t) { int64_t b; memcpy(&b, &t, 8); return (int)(uint32_t)b; }
GEOAC_HD double g_rcp(double x)   { return 1.0 / x; }
GEOAC_HD double g_rsqrt(double x) { return 1.0 / sqrt(x); }
GEOAC_HD double g_sqrt_rs(double x, double& rs) { const double s = sqrt(x); rs = 1.0 / s; return s; }
GEOAC_HD double g_rcbrt(double x) { return 1.0 / cbrt(x); }
#endif
GEOAC_HD double g_sqrt(double x) { double rs; return g_sqrt_rs(x, rs); }

// N exponentials in lock step (explicit ILP): out[j] = exp(x[j]) (BASE10 = false) or 10^x[j] (BASE10 = true),
// for |result exponent| < 1020: the exponent is patched directly, arguments must stay inside +-700 (e) / +-300 (10).
template <bool BASE10, int N>
GEOAC_HD void g_exp_n(const double (&x)[N], double (&out)[N]) {     // caller guarantees |x| < 700 (e) / 300 (10)
    const double MAGIC = 6755399441055744.0;                                  // 1.5 * 2^52: rint() by addition
    const double L2B = BASE10 ? 3.321928094887362 : 1.4426950408889634;       // log2(base)
    const double HI = BASE10 ? 0.3010299956639812 : 0.6931471805599453;       // log_base(2), split
    const double LO = BASE10 ? -2.8037281277851704e-18 : 2.3190468138462996e-17;
#ifdef GEOAC_COUNT_FLOPS      // tests/flopcount: one libm call = one transcendental, not its polynomial expansion
#pragma unroll
    for (int j = 0; j < N; j++) out[j] = BASE10 ? pow(10.0, x[j]) : exp(x[j]);
    return;
#endif
    const double* C = BASE10 ? kExp10 : kExpE;
    double r[N], p[N]; int n[N];
#pragma unroll
    for (int j = 0; j < N; j++) {
        const double xc = x[j];
        const double t = fma(xc, L2B, MAGIC);
        n[j] = g_lo32(t);
        const double nf = t - MAGIC;
        r[j] = fma(nf, -LO, fma(nf, -HI, xc));
        p[j] = C[11];
    }
#pragma unroll
    for (int k = 10; k >= 0; k--) {
#pragma unroll
        for (int j = 0; j < N; j++) p[j] = fma(p[j], r[j], C[k]);
    }
#pragma unroll
    for (int j = 0; j < N; j++) out[j] = g_scale2(p[j], n[j]);
}
// sin and cos together for |x| up to a few hundred (latitudes, half-angle differences): Cody-Waite reduction by pi/2 with
// the FMA, the classic minimax kernels on |r| <= pi/4 (error < 2^-58), quadrant fix-up by selects -- no slow path.
GEOAC_CONST_TABLE double kSinC[6] = { -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
                                      2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10 };
GEOAC_CONST_TABLE double kCosC[6] = { 4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
                                      -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11 };
GEOAC_HD void g_sincos(double x, double* s, double* c) {
#if defined(__CUDA_ARCH__)
    const double MAGIC = 6755399441055744.0;
    const double t = fma(x, 0.6366197723675814, MAGIC);                      // x * 2/pi, rounded to an integer by addition
    const int q = g_lo32(t);
    const double k = t - MAGIC;
    double r = fma(-k, 1.5707963267948966, x);
    r = fma(-k, 6.123233995736766e-17, r);
    const double z = r * r;
    double ps = kSinC[5], pc = kCosC[5];
#pragma unroll
    for (int i = 4; i >= 0; i--) { ps = fma(ps, z, kSinC[i]); pc = fma(pc, z, kCosC[i]); }
    const double sr = fma(r * z, ps, r);
    const double cr = fma(z * z, pc, fma(-0.5, z, 1.0));
    const double ss = (q & 1) ? cr : sr, cc = (q & 1) ? sr : cr;
    *s = (q & 2) ? -ss : ss;
    *c = ((q + 1) & 2) ? -cc : cc;
#else
    sincos(x, s, c);
#endif
}

GEOAC_HD double g_exp(double x) { const double a[1] = { x }; double o[1]; g_exp_n<false, 1>(a, o); return o[0]; }

// ---------------------------------------------------------------------------------------------------------------
// 1-D table: one 80-byte record per level,
//   { x, invh, T, sT, u, su, v, sv, rho, srho }          (invh[k] = 1/(x[k+1]-x[k]), invh[n-1] = 0; s* = knot slopes)
// so that a sample touches two adjacent records and every (value, slope) pair -- and (x, invh) -- is one aligned 16-byte
// load: 8 LDS.128 per RK4 stage instead of 15 LDS.64, with one address computation per sample.
// ---------------------------------------------------------------------------------------------------------------
enum { TAB_X = 0, TAB_INVH, TAB_T, TAB_ST, TAB_U, TAB_SU, TAB_V, TAB_SV, TAB_RHO, TAB_SRHO, TAB_NARR };

struct Table1D {
    const double* base;     // shared or global memory, 16-byte aligned on the device
    int n;
    double xmin, xmax;
    double jump_scale;      // > 0 (cost scout only): (n-1)/(xmax-xmin), lets a far query jump near its interval before the walk
    const double* sbpoly = nullptr;   // per-interval absorption polynomials (SBP_STRIDE doubles each, global memory) or nullptr
    GEOAC_HD const double* lvl(int k) const { return base + (size_t)k * TAB_NARR; }
};

#if defined(__CUDA_ARCH__)
struct Pair { double a, b; };
GEOAC_HD Pair ld_pair(const double* p) { const double2 v = *reinterpret_cast<const double2*>(p); Pair r; r.a = v.x; r.b = v.y; return r; }
#else
struct Pair { double a, b; };
GEOAC_HD Pair ld_pair(const double* p) { Pair r; r.a = p[0]; r.b = p[1]; return r; }
#endif

struct SegPos { const double* r0; double X, h, invh; };      // r0 = record of the interval's lower level (upper: r0 + TAB_NARR)

// clamp without the NaN plumbing of fmin/fmax (one compare + select per bound)
GEOAC_HD double clampd(double v, double lo, double hi) { v = (v > hi) ? hi : v; return (v < lo) ? lo : v; }

// BREAK margin (GEOAC_F_MARGIN of a BREAK slot): fraction of the step a -> b that lay beyond a limit, from the excesses
// e = value - limit (> 0 outside) at both ends; the smallest over the violated limits is kept
GEOAC_HD double frac_beyond(double ea, double eb, double best) { return (eb > 0.0) ? fmin(best, eb / (eb - ea)) : best; }

// locate the interval containing x, clamped into the table range like every reference look-up (G2S_Spline1D.cpp:335),
// starting from cursor k.  Same tie-breaking as the reference's Find_Segment (G2S_Spline1D.cpp:202-243): a point on a
// knot stays in the interval the cursor is already in.  Fast path: the query is still inside the cursor's interval
// (then it needs no clamping either) -- one 16-byte and one 8-byte load, two compares.
GEOAC_HD SegPos seg_locate(const Table1D& t, double xq, int& k) {
    const double* r = t.lvl(k);
    Pair lo = ld_pair(r);                                     // x[k], invh[k]
    double x1 = r[TAB_NARR];
    double xc = xq;
    if (!(xq >= lo.a && xq <= x1)) {
        xc = clampd(xq, t.xmin, t.xmax);
        if (t.jump_scale > 0.0) {            // the scout strides many levels per stage; knot tie-breaking does not matter to it
            int kg = (int)((xc - t.xmin) * t.jump_scale);
            kg = (kg < 0) ? 0 : ((kg > t.n - 2) ? t.n - 2 : kg);
            k = kg; r = t.lvl(k); lo = ld_pair(r); x1 = r[TAB_NARR];
        }
        while (xc < lo.a) { --k; r -= TAB_NARR; x1 = lo.a; lo = ld_pair(r); }
        while (xc > x1)   { ++k; r += TAB_NARR; lo = ld_pair(r); x1 = r[TAB_NARR]; }
    }
    SegPos s; s.r0 = r; s.h = x1 - lo.a; s.invh = lo.b; s.X = (xc - lo.a) * lo.b;
    return s;
}

// Hermite "slopes" form of the natural cubic spline on one interval (G2S_Spline1D.cpp:245-281), arranged for FMAs:
//   A = s0 h - df, B = df - s1 h, C = B - A, P = A + C X
//   f = f0 + X df + X(1-X) P,   f' h = df + (1-2X) P + X(1-X) C,   f'' h^2 = 2 ((1-2X) C - P)
struct SplCoef { double f0, df, A, C; };
GEOAC_HD SplCoef spl_coef(const SegPos& p, int field) {
    const Pair a = ld_pair(p.r0 + field), b = ld_pair(p.r0 + TAB_NARR + field);       // (f, slope) at both levels
    SplCoef c; c.f0 = a.a; c.df = b.a - a.a;
    c.A = fma(a.b, p.h, -c.df);
    c.C = fma(-b.b, p.h, c.df) - c.A;
    return c;
}
// value only
GEOAC_HD double spl_f(const Table1D&, int field, const SegPos& p) {
    const SplCoef c = spl_coef(p, field);
    const double XomX = p.X * (1.0 - p.X);
    return fma(XomX, fma(c.C, p.X, c.A), fma(p.X, c.df, c.f0));
}
// value + first derivative
GEOAC_HD void spl_f1(const Table1D&, int field, const SegPos& p, double& f, double& d1) {
    const SplCoef c = spl_coef(p, field);
    const double XomX = p.X * (1.0 - p.X), om2X = fma(-2.0, p.X, 1.0), P = fma(c.C, p.X, c.A);
    f  = fma(XomX, P, fma(p.X, c.df, c.f0));
    d1 = fma(XomX, c.C, fma(om2X, P, c.df)) * p.invh;
}
// value + first + second derivative
GEOAC_HD void spl_f2(const Table1D&, int field, const SegPos& p, double& f, double& d1, double& d2) {
    const SplCoef c = spl_coef(p, field);
    const double XomX = p.X * (1.0 - p.X), om2X = fma(-2.0, p.X, 1.0), P = fma(c.C, p.X, c.A);
    f  = fma(XomX, P, fma(p.X, c.df, c.f0));
    d1 = fma(XomX, c.C, fma(om2X, P, c.df)) * p.invh;
    d2 = fma(om2X, c.C, -P) * (2.0 * p.invh * p.invh);
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
struct SBRef { double invTo, cbrtTo, visc_num, inv_visc_num, invPo; };

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
    // log10 gas-fraction polynomials with an altitude branch, [0] = low / [1] = high (Absorption.cpp:70-99): read with a
    // lane-dependent branch index, so they sit in shared memory next to the other invariants rather than in constant memory
    double sbx3[2][6], sbx4[2][4], sbx6[2][6];
};

// GeoAc_Set_ds (3DStratified.cpp:191-198 and siblings): ds = clamp(0.05 - 0.049 exp(-h / 0.75), ds_min, ds_max), h = height above
// ground.  Above 30 km the exponential term is below 2^-58, half an ulp of 0.05, so the rounded difference IS 0.05: the
// polynomial is skipped there (a warp's rays have near-equal altitudes, so the branch rarely diverges) without changing a bit.
GEOAC_HD double step_size_z(const LaunchConsts& L, double h) {
    double r = 0.05;
    if (h < 30.0) r = 0.05 - 0.049 * g_exp(-h * (1.0 / 0.75));
    return fmax(fmin(r, L.ds_max), L.ds_min);
}

// log10 of the gas fractions as polynomials in altitude [km] (Atmo_State.Absorption.cpp:55-99); [0] = low branch, [1] = high branch
GEOAC_CONST_TABLE double kSBX0[6] = { 49.296, -1.5524, 1.8714E-2, -1.1069E-4, 3.199E-7, -3.6211E-10 };                    // O2, z > 90
GEOAC_CONST_TABLE double kSBX1[4] = { 1.3972E-1, -5.6269E-3, 3.9407E-5, -1.0737E-7 };                                       // N2, z > 76
GEOAC_CONST_TABLE double kSBX5[6] = { -53.746, 1.5439, -1.8824E-2, 1.1587E-4, -3.5399E-7, 4.2609E-10 };                   // N

// Sutherland-Bass absorption [dB/km] at altitude z [km] with local sound speed c [km/s] (and 1/c) and density rho.
// Same model and branch thresholds (strict >) as Atmo_State.Absorption.cpp:14-143, restructured for the FP64 pipe:
//   * the ten exp(a_i Tr) factors of the vibrational relaxation frequencies share ONE exponential: every a_i is a
//     multiple of 0.01, so they are integer powers of g = exp(-Tr/100) (binary powers, ~40 multiplies; the relative
//     error grows to <= 2000 * 2^-53 ~ 2e-13, four orders inside the parity budget);
//   * the remaining exponentials (gas fractions 10^poly(z), rotational collision numbers, vibrational Boltzmann factors)
//     are evaluated in lock step by the branch-free g_exp_n; the polynomials in z are Horner forms;
//   * every quotient is a product with one of five reciprocals (two of them batched inversions).
// `parts` (table builder only): [0] = G2 with (a_cl + a_diff) * scale = sqrt(s1m1 * G2), [1] = (a_rot + a_vib) * scale, [2] = nu^2,
// where scale = tweak_abs * 8.685889 and s1m1 = sqrt(1 + nu^2) - 1 is the one non-smooth factor of the model (see below).
GEOAC_HD double suthbass_alpha(const LaunchConsts& L, const SBRef& R, double z, double c, double inv_c, double rho, double* parts = nullptr) {
    const double mu_o = 18.192E-6, S = 117.0;
    const double inv_c2 = inv_c * inv_c;
    const double c2 = (c * c) * 1.0e6;                                      // (1000 c)^2
    const double T_z = c2 * (1.0 / (kR * kGam));
    const double inv_Tz = inv_c2 * (kR * kGam * 1.0e-6);
    const double P_z = rho * c2 * (1000.0 / kGam);
    const double den1 = 1.0 + S * inv_Tz;
    const double r1 = g_rcp(rho * den1);
    const double inv_rho = r1 * den1, inv_den1 = r1 * rho;
    const double inv_Pz = inv_rho * inv_c2 * (kGam * 1.0e-9);

    const double tq = T_z * R.invTo;
    double rs; const double sq = g_sqrt_rs(tq, rs);
    const double mu_ratio = sq * (R.visc_num * inv_den1);                    // mu/mu_o
    const double inv_mu_ratio = rs * (den1 * R.inv_visc_num);
    const double nu = (8.0 * kPi * L.freq * mu_o * (1.0 / 3.0)) * mu_ratio * inv_Pz;

    // gas fractions X0..X6 = 10^(polynomial in z), Absorption.cpp:55-99.  Coefficient sets live in constant memory and are
    // picked by the branch predicate, so the warp never diverges on the altitude thresholds; z is clamped to the range in
    // which every polynomial stays representable (the reference under/overflows to 0 / inf beyond it).
    const double zc = fmax(fmin(z, 200.0), -20.0);
    double X0 = 0.20947393930547747;                                       // 10^-0.67887
    double X1 = 0.780836309209143;                                         // 10^-0.10744
    const double X2 = 0.00040003685104612505;                              // 10^-3.3979
    if (z > 76.) {
        const double q[2] = { kSBX0[0] + zc * (kSBX0[1] + zc * (kSBX0[2] + zc * (kSBX0[3] + zc * (kSBX0[4] + zc * kSBX0[5])))),
                              kSBX1[0] + zc * (kSBX1[1] + zc * (kSBX1[2] + zc * kSBX1[3])) };
        double e[2]; g_exp_n<true, 2>(q, e);
        if (z > 90.) X0 = e[0];
        X1 = e[1];
    }
    double X3, X4, X5, X6;
    {
        const double* a = L.sbx3[z > 80. ? 1 : 0];
        const double* bq = L.sbx4[z > 95. ? 1 : 0];
        const double* d = L.sbx6[z > 30. ? 1 : 0];
        const double q[4] = { a[0] + zc * (a[1] + zc * (a[2] + zc * (a[3] + zc * (a[4] + zc * a[5])))),
                              bq[0] + zc * (bq[1] + zc * (bq[2] + zc * bq[3])),
                              kSBX5[0] + zc * (kSBX5[1] + zc * (kSBX5[2] + zc * (kSBX5[3] + zc * (kSBX5[4] + zc * kSBX5[5])))),
                              d[0] + zc * (d[1] + zc * (d[2] + zc * (d[3] + zc * (d[4] + zc * d[5])))) };
        double e[4]; g_exp_n<true, 4>(q, e);
        X3 = e[0]; X4 = e[1]; X5 = e[2]; X6 = e[3];
    }
    const double X01 = X0 + X1;
    const double X_ON = X01 * (1.0 / 0.9903);

    // natural exponentials: g = exp(-Tr/100); rotational collision numbers; vibrational Boltzmann factors
    const double cb = g_rcbrt(T_z);                                        // T_z^(-1/3)
    const double Tr = cb * R.cbrtTo - 1.0;                                 // (T_z/T_o)^(-1/3) - 1
    const double th[4] = { 2239.1, 3352.0, 915.0, 1037.0 };
    double ex[7];
    {
        const double q[7] = { -0.01 * Tr, -17.3 * cb, -16.7 * cb, -th[0] * inv_Tz, -th[1] * inv_Tz, -th[2] * inv_Tz, -th[3] * inv_Tz };
        g_exp_n<false, 7>(q, ex);
    }
    const double Zr0 = 54.1 * ex[1], Zr1 = 63.3 * ex[2];
    const double Z_rot_ = (Zr0 * Zr1) * g_rcp(X1 * Zr0 + X0 * Zr1);         // 1/(X1/Zr1 + X0/Zr0)

    const double sigma = 1.091089451179962;                                // 5/sqrt(21)
    const double nn = 0.5237229365663817 * Z_rot_;                         // (4/5) sqrt(3/7) Z_rot_
    const double chi = 0.75 * nn * nu;
    const double cchi = 2.36 * chi;
    const double nu2p1 = 1.0 + nu * nu;
    const double s1 = g_sqrt(nu2p1);
    const double cchi2p1 = 1.0 + cchi * cchi;
    const double sc = sigma * cchi;
    const double w_c = L.sb_w * inv_c;                                     // 2 pi f / c
    // NB: sqrt(1+nu^2) - 1 cancels catastrophically for small nu (it is 0 or a few ulp below ~60 km at 0.1 Hz).
    // That quantisation IS the reference's observable behaviour (Absorption.cpp:108), so it is reproduced literally.
    const double s1m1 = s1 - 1.0;
    const double q1 = nu2p1 * (1.0 + sc * sc), q2 = nu2p1 * cchi2p1;
    const double rq = g_rcp(q1 * q2);                                      // 1/q1 = rq q2, 1/q2 = rq q1
    const double a_cl  = w_c * g_sqrt(fmax(0.5 * s1m1 * cchi2p1 * (rq * q2), 1e-290));
    const double a_rot = w_c * X_ON * ((sigma * sigma - 1.0) * chi * (0.5 / sigma)) * g_sqrt(0.5 * (s1 + 1.0) * (rq * q1));
    const double a_diff = 0.003 * a_cl;

    // integer powers of g (binary exponentiation): g2 = g^2, g4 = g^4, ... g1024
    const double g1 = ex[0], g2 = g1 * g1, g4 = g2 * g2, g8 = g4 * g4, g16 = g8 * g8, g32 = g16 * g16, g64 = g32 * g32,
                 g128 = g64 * g64, g256 = g128 * g128, g512 = g256 * g256, g1024 = g512 * g512;
    const double g768 = g512 * g256, g896 = g768 * g128;
    const double g916 = g896 * (g16 * g4);                                 // exp(-9.16 Tr)
    const double g917 = g916 * g1;                                         // exp(-9.17 Tr)
    const double g1120 = g1024 * (g64 * g32);                              // exp(-11.2 Tr)
    const double g1990 = (g1024 * g896) * ((g64 * g4) * g2);               // exp(-19.9 Tr)
    const double g417 = (g256 * g128) * (g32 * g1);                        // exp(-4.17 Tr)
    const double g1040 = g1024 * g16;                                      // exp(-10.4 Tr)
    const double g772 = g768 * g4;                                         // exp(-7.72 Tr)
    const double g1000 = g896 * ((g64 * g32) * g8);                        // 1/exp(10 Tr)
    const double g841 = g768 * ((g64 * g8) * g1);                          // 1/exp(8.41 Tr)
    // batched inversion of g1000, g841, g917
    const double m12 = g1000 * g841;
    const double rall = g_rcp(m12 * g917);
    const double i917 = rall * m12, r12 = rall * g917, i1000 = r12 * g841, i841 = r12 * g1000;

    const double A1 = X01 * 24.0 * g916;
    const double A2 = (X4 + X5) * 2400.0;
    const double B  = 40400.0 * i1000;
    const double C  = 0.02 * g1120;
    const double D  = 0.391 * i841;
    const double E  = 9.0 * g1990;
    const double F  = 60000.0;
    const double G  = 28000.0 * g417;
    const double H  = 22000.0 * g768;
    const double I  = 15100.0 * g1040;
    const double J  = 11500.0 * g917;
    const double K  = (8.48E08) * i917;
    const double ZZ = H * X2 + I * (X0 + 0.5 * X4) + J * (X1 + 0.5 * X5) + K * (X6 + X3);
    const double hu = 100.0 * (X3 + X6);
    const double pm = (P_z * R.invPo) * inv_mu_ratio;                       // (P_z/P_o)(mu_o/mu)
    double fv[4];
    fv[0] = pm * (A1 + A2 + B * hu * (C + hu) * (D + hu));
    fv[1] = pm * (E + F * X3 + G * X6);
    fv[2] = pm * ZZ;
    fv[3] = pm * (1.2E5) * g772;

    // vibrational terms: A_max/c * (2 f^2/fv)/(1 + (f/fv)^2) with A_max = X (pi/2) C_R / (Cp (Cv + C_R)), C_R = r^2 e/(1-e)^2
    //   = X (pi/2) r^2 e * 2 f^2 fv / ( c * Cp (Cv (1-e)^2 + r^2 e) (fv^2 + f^2) ): one quotient per term, inverted together
    const double CpR[4] = { 3.5, 3.5, 4.0, 4.0 }, CvR[4] = { 2.5, 2.5, 3.0, 3.0 };
    const double Xm[4] = { X0, X1, X2, X3 };
    const double f2 = L.freq * L.freq;
    double num[4], den[4];
#pragma unroll
    for (int m = 0; m < 4; m++) {
        const double r = th[m] * inv_Tz;
        const double e = ex[3 + m];
        const double ome = 1.0 - e;
        const double r2e = (r * r) * e;
        num[m] = (Xm[m] * (kPi / 2) * 2.0) * r2e * (f2 * fv[m]);
        den[m] = CpR[m] * (CvR[m] * (ome * ome) + r2e) * (fv[m] * fv[m] + f2);
    }
    const double d01 = den[0] * den[1], d23 = den[2] * den[3];
    const double rd = g_rcp(d01 * d23);
    const double r01 = rd * d23, r23 = rd * d01;
    const double a_vib = inv_c * ((num[0] * (r01 * den[1]) + num[1] * (r01 * den[0])) + (num[2] * (r23 * den[3]) + num[3] * (r23 * den[2])));
    if (parts) {
        const double sc2 = 1.003 * L.tweak_abs * 8.685889;
        parts[0] = (w_c * w_c) * (0.5 * cchi2p1 * (rq * q2)) * (sc2 * sc2);
        parts[1] = (a_rot + a_vib) * L.tweak_abs * 8.685889;
        parts[2] = nu * nu;
    }
    return (a_cl + a_rot + a_diff + a_vib) * L.tweak_abs * 8.685889;
}

// ---- absorption of the stratified variants through per-interval polynomials ----
// In a stratified atmosphere c and rho are functions of altitude alone, so the Sutherland-Bass coefficient is too, and inside
// one spline interval (0.1 km in the G2S profiles) it is analytic EXCEPT for the classical term's factor
// s1m1 = sqrt(1 + nu^2) - 1, whose cancellation noise (a staircase in z below ~60 km) is observable in the reference's output
// and is therefore kept literal.  alpha = sqrt(s1m1 * G2(z)) + S(z): nu and s1m1 are computed exactly as in suthbass_alpha,
// G2 and S -- all the exponentials -- are degree-6 interpolants of the exact function at the interval's Chebyshev nodes,
// built once per launch configuration (freq, abs_coeff) by sbpoly_build_interval with the device's own suthbass_alpha.
// An interval whose interpolants miss the exact function by more than 1e-12 relative at eight check points (a gas-fraction
// threshold inside it, a coarse user profile) is flagged and evaluated exactly, as are queries outside the table.
constexpr int SBP_DEG = 6;
constexpr int SBP_STRIDE = 24;      // [0..6] G2 (monomials in s = 2X - 1), [7] flag (0 = valid), [8..14] S, [16..22] nu^2, [15], [23] unused

#if defined(__CUDA_ARCH__)
GEOAC_HD Pair ldg_pair(const double* p) { const double2 v = __ldg(reinterpret_cast<const double2*>(p)); Pair r; r.a = v.x; r.b = v.y; return r; }
#else
GEOAC_HD Pair ldg_pair(const double* p) { Pair r; r.a = p[0]; r.b = p[1]; return r; }
#endif

// the full model, out of line: it is the rare path of sb_alpha_1d and must not set the register budget of the step loop
#if defined(__CUDA_ARCH__)
__device__ __noinline__ double suthbass_alpha_cold(const LaunchConsts& L, double z, double c, double inv_c, double rho) { return suthbass_alpha(L, L.sb, z, c, inv_c, rho); }
#else
inline double suthbass_alpha_cold(const LaunchConsts& L, double z, double c, double inv_c, double rho) { return suthbass_alpha(L, L.sb, z, c, inv_c, rho); }
#endif

// absorption [dB/km] at the point seg_locate(T, zq, k) found (interval k, offset sp.X); z_eff = altitude the model sees.
// nu^2 (the viscous-to-pressure ratio squared: a smooth function of c and rho, hence of z) comes from its own interpolant; the
// cancelling difference sqrt(1 + nu^2) - 1 is then formed exactly as the reference forms it, so its quantisation staircase
// stays (a rounding-level change of nu^2 moves a step of the staircase by the same relative amount, nothing else).
GEOAC_HD double sb_alpha_1d(const LaunchConsts& L, const Table1D& T, const SegPos& sp, int k, double zq, double z_eff, double c, double inv_c) {
    if (T.sbpoly != nullptr && zq >= T.xmin && zq <= T.xmax) {
        const double* q = T.sbpoly + (size_t)k * SBP_STRIDE;
        const Pair g67 = ldg_pair(q + 6);
        if (g67.b == 0.0) {
            const Pair g01 = ldg_pair(q), g23 = ldg_pair(q + 2), g45 = ldg_pair(q + 4);
            const Pair s01 = ldg_pair(q + 8), s23 = ldg_pair(q + 10), s45 = ldg_pair(q + 12), s67 = ldg_pair(q + 14);
            const Pair n01 = ldg_pair(q + 16), n23 = ldg_pair(q + 18), n45 = ldg_pair(q + 20), n67 = ldg_pair(q + 22);
            const double s = fma(2.0, sp.X, -1.0);
            const double nu2 = fma(fma(fma(fma(fma(fma(n67.a, s, n45.b), s, n45.a), s, n23.b), s, n23.a), s, n01.b), s, n01.a);
            const double G2 = fma(fma(fma(fma(fma(fma(g67.a, s, g45.b), s, g45.a), s, g23.b), s, g23.a), s, g01.b), s, g01.a);
            const double Sm = fma(fma(fma(fma(fma(fma(s67.a, s, s45.b), s, s45.a), s, s23.b), s, s23.a), s, s01.b), s, s01.a);
            const double s1m1 = g_sqrt(1.0 + nu2) - 1.0;
            return g_sqrt(fmax(s1m1 * G2, 1e-290)) + Sm;
        }
    }
    return suthbass_alpha_cold(L, z_eff, c, inv_c, spl_f(T, TAB_RHO, sp));
}

// Builds the SBP_STRIDE coefficients of interval k of a 1-D table (one thread per interval on the device).
GEOAC_HD void sbpoly_build_interval(const LaunchConsts& L, const Table1D& T, bool glob, int k, double* out) {
    const double x0 = T.lvl(k)[TAB_X], x1 = T.lvl(k + 1)[TAB_X], h = x1 - x0;
    double f[3][SBP_DEG + 1], a[3][SBP_DEG + 1];
    auto eval = [&](double s, double* g2, double* sm, double* n2) {
        double z = x0 + (0.5 * (s + 1.0)) * h;
        z = (z < x0) ? x0 : ((z > x1) ? x1 : z);
        int cur = k;
        const SegPos sp = seg_locate(T, z, cur);
        const double Tv = spl_f(T, TAB_T, sp), rho = spl_f(T, TAB_RHO, sp);
        const double gT = kGamR * Tv;
        const double inv_c = g_rsqrt(gT), c = gT * inv_c;
        double parts[3];
        suthbass_alpha(L, L.sb, glob ? z - kREarth : z, c, inv_c, rho, parts);
        *g2 = parts[0]; *sm = parts[1]; *n2 = parts[2];
    };
    const int N = SBP_DEG + 1;
    for (int j = 0; j < N; j++) eval(cos((2 * j + 1) * (kPi / (2.0 * N))), &f[0][j], &f[1][j], &f[2][j]);
    for (int w = 0; w < 3; w++) {
        for (int m = 0; m < N; m++) {
            double acc = 0.0;
            for (int j = 0; j < N; j++) acc += f[w][j] * cos((double)(m * (2 * j + 1)) * (kPi / (2.0 * N)));
            a[w][m] = acc * ((m == 0 ? 1.0 : 2.0) / N);
        }
        double* o = out + 8 * w;                                           // Chebyshev -> monomials (degree 6)
        o[0] = a[w][0] - a[w][2] + a[w][4] - a[w][6];
        o[1] = a[w][1] - 3.0 * a[w][3] + 5.0 * a[w][5];
        o[2] = 2.0 * a[w][2] - 8.0 * a[w][4] + 18.0 * a[w][6];
        o[3] = 4.0 * a[w][3] - 20.0 * a[w][5];
        o[4] = 8.0 * a[w][4] - 48.0 * a[w][6];
        o[5] = 16.0 * a[w][5];
        o[6] = 32.0 * a[w][6];
    }
    bool bad = false;
    const double chk[8] = { -0.97, -0.75, -0.45, -0.15, 0.15, 0.45, 0.75, 0.97 };
    for (int j = 0; j < 8; j++) {
        double e[3]; eval(chk[j], &e[0], &e[1], &e[2]);
        for (int w = 0; w < 3; w++) {
            const double* o = out + 8 * w; const double s = chk[j];
            const double pv = o[0] + s * (o[1] + s * (o[2] + s * (o[3] + s * (o[4] + s * (o[5] + s * o[6])))));
            if (!(fabs(pv - e[w]) <= 1e-12 * fabs(e[w])) || !(e[w] > 0.0)) bad = true;
        }
    }
    out[7] = bad ? 1.0 : 0.0; out[15] = 0.0; out[23] = 0.0;
}

// fill the Sutherland-Bass invariants from the reference state (c, rho at the reference level)
GEOAC_HD void suthbass_ref(SBRef& R, double c_ref, double rho_ref) {
    const double c1000 = c_ref * 1000.0;
    const double T_o = c1000 * c1000 / (kR * kGam);
    const double P_o = rho_ref * (c1000 * c1000) / kGam * 1000.0;
    R.invTo = 1.0 / T_o; R.cbrtTo = cbrt(T_o); R.visc_num = 1.0 + 117.0 / T_o; R.inv_visc_num = 1.0 / R.visc_num;
    R.invPo = 1.0 / P_o;
}
GEOAC_HD void suthbass_tables(LaunchConsts& L) {
    const double x3[2][6] = { { -19.027, 1.3093, -4.6496E-2, 7.8543E-4, -6.5169E-6, 2.1343E-8 }, { -4.234, -3.0975E-2, 0.0, 0.0, 0.0, 0.0 } };   // O3: z <= 80, z > 80
    const double x4[2][4] = { { -11.195, 1.5408E-1, -1.4348E-3, 1.0166E-5 }, { -3.2456, 4.6642E-2, -2.6894E-4, 5.264E-7 } };                     // O:  z <= 95, z > 95
    const double x6[2][6] = { { -1.7491, 4.4986E-2, -6.8549E-2, 5.4639E-3, -1.5539E-4, 1.5063E-06 },                                            // H2O: z <= 30
                              { -4.2563, 7.6245E-2, -2.1824E-3, -2.3010E-6, 2.4265E-7, -1.2500E-09 } };                                         //      z > 30
    for (int b = 0; b < 2; b++) {
        for (int k = 0; k < 6; k++) { L.sbx3[b][k] = x3[b][k]; L.sbx6[b][k] = x6[b][k]; }
        for (int k = 0; k < 4; k++) L.sbx4[b][k] = x4[b][k];
    }
    L.sb_w = 2.0 * kPi * L.freq;
}
GEOAC_HD void suthbass_setup(LaunchConsts& L, double c_ref, double rho_ref) {
    suthbass_ref(L.sb, c_ref, rho_ref);
    suthbass_tables(L);
}

}  // namespace geoac
