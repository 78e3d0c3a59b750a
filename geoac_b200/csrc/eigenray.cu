// eigenray.cu -- batched eigenray search above the C ABI (host code only; every ray goes through geoac_trace on the GPU).
//
// Replaces the body of GeoAc3D_RunEigSearch (Code/GeoAc3D_main.cpp:531-541, Code/GeoAc3D.RngDep_main.cpp likewise) and of
// GeoAcGlobal_RunEigSearch (Code/GeoAcGlobal_main.cpp:573-583, GeoAcGlobal.RngDep_main.cpp likewise; the Global routines
// are Code/GeoAc/GeoAc.Eigenray.Global.cpp:46-136 and :139-320 -- same algorithm on great-circle range and bearing):
//     for every bounce count:  theta_start = theta_min;
//         while theta_start < theta_max:  GeoAc_EstimateEigenray (Code/GeoAc/GeoAc.Eigenray.cpp:30-121)
//                                         on success GeoAc_3DEigenray_LM (:123-335);  theta_start = theta_next
// The reference traces one ray at a time (a 0.25-degree inclination fan per estimate, one ray per Levenberg-Marquardt
// iteration).  Here the two routines are restated literally over a RAY CACHE keyed by the exact launch angles: a look-up
// that misses records a request -- for the fixed-step fans of the estimate, the whole remaining fan -- and the search is
// replayed from the start once the requests of ALL receivers, bounce counts and brackets have been traced in ONE
// geoac_trace batch.  A replay only ever reads rays the sequential algorithm would have traced at exactly those angles, so
// the final replay IS the reference's sequence of decisions; batching changes only when rays are computed.
//   * the iteration-0 fan (one azimuth, theta_min..theta_max) is shared by every estimate call and every bounce count of a
//     receiver: a trace with `bounces = B` yields the arrivals of all bounce counts <= B;
//   * later estimate calls of a chain start from theta_next, which iteration 0 already fixes, so all brackets of a fan
//     refine their azimuth and run their LM iterations concurrently (speculatively: a call that ends on the reference's
//     theta_max_reached quirk cuts the chain, and the final replay honours that);
//   * LM iterations are one ray per live bracket per round (sequentially dependent inside a bracket, as in the reference).
// `long double` is used exactly where the reference uses it (x87 extended on x86-64).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <map>
#include <string>
#include <tuple>
#include <vector>
#include "../../include/geoac_b200.h"

namespace {

constexpr double Pi = 3.141592653589793238462643;      // Code/GeoAc/GeoAc.Parameters.cpp:28

struct RayKey {
    uint64_t tb, pb; int amp;
    bool operator<(const RayKey& o) const { return std::tie(tb, pb, amp) < std::tie(o.tb, o.pb, o.amp); }
};
struct RayVal {
    int nb = -1;                              // bounces traced (records 0..nb)
    std::vector<int32_t> status;              // [nb+1]
    std::vector<double> state;                // [nb+1][18]
};
static uint64_t bits(double x) { uint64_t u; std::memcpy(&u, &x, 8); return u; }

struct Request { double theta, phi; int amp, nb; };

struct Search {
    geoac_ctx* ctx;
    int variant;
    bool strat, glob = false;
    geoac_eig_opts o;
    double src[3];                            // Source_Loc: (x, y, z) [km], or (lat, lon) [deg] and z for the Global variants
    double z_grnd = 0.0;
    double Mx = 0, My = 0;                    // u/c, v/c at the source (stratified LM, Eigenray.cpp:131-135; w = 0)
    std::map<RayKey, RayVal> cache;
    std::map<RayKey, Request> want;           // this round's misses
    int64_t rays_traced = 0; int rounds = 0;

    // state of the ray launched at (GeoAc_theta, GeoAc_phi) [radians, exactly the values the reference assigns] after n_bnc
    // reflections, or nullptr + a recorded request.  *brk receives the reference's BreakCheck (any segment left the region;
    // a step-limit end counts as one, see include/geoac_b200.h).
    const double* lookup(double theta, double phi, int amp, int n_bnc, bool* brk, bool record = true) {
        RayKey k{ bits(theta), bits(phi), amp };
        auto it = cache.find(k);
        if (it != cache.end()) {
            const RayVal& v = it->second;
            bool ended = false;
            for (int b = 0; b <= std::min(n_bnc, v.nb); b++) if (v.status[b] != GEOAC_ST_ARRIVAL) { ended = true; break; }
            if (ended) { *brk = true; return v.state.data(); }
            if (v.nb >= n_bnc) { *brk = false; return v.state.data() + (size_t)n_bnc * 18; }
        }
        if (record) {
            auto w = want.find(k);
            if (w == want.end()) want[k] = Request{ theta, phi, amp, n_bnc };
            else w->second.nb = std::max(w->second.nb, n_bnc);
        }
        return nullptr;
    }

    // ---- GeoAc_EstimateEigenray, Eigenray.cpp:30-121 ----
    struct Est { bool complete = false, ok = false, next_known = false; double theta_est = 0, phi_est = 0, theta_next = 0; };
    static double modify_d_theta(double dr, double dr_dtheta, double big, double small_, double wf) {   // :23-27; Global :39-43 (width factor 1/2)
        const double width = wf * std::pow(dr_dtheta, 2);
        return big - (big - small_) * std::exp(-dr * dr / width);
    }
    static double bearing(double lat1, double long1, double lat2, double long2) {                 // Calc_Bearing, Eigenray.Global.cpp:25-30
        const double term1 = std::sin((long2 - long1) * Pi / 180.0);
        const double term2 = std::cos(lat1 * Pi / 180.0) * std::tan(lat2 * Pi / 180.0) - std::sin(lat1 * Pi / 180.0) * std::cos((long2 - long1) * Pi / 180.0);
        return std::atan2(term1, term2) * 180.0 / Pi;
    }
    static double gc_distance(double lat1, double long1, double lat2, double long2) {             // Calc_GC_Distance, :32-37
        const double term1 = std::pow(std::sin((lat2 - lat1) * Pi / 180.0 / 2.0), 2);
        const double term2 = std::cos(lat1 * Pi / 180.0) * std::cos(lat2 * Pi / 180.0) * std::pow(std::sin((long2 - long1) * Pi / 180.0 / 2.0), 2);
        return 2.0 * 6370.0 * std::asin(std::sqrt(term1 + term2));
    }
    Est estimate(const double rcv[2], double theta_min, double theta_max, int bounces) {
        Est e;
        const double r_rcvr = glob ? gc_distance(src[0], src[1], rcv[0], rcv[1])
                                   : std::sqrt(std::pow(rcv[0] - src[0], 2) + std::pow(rcv[1] - src[1], 2));
        double phi = glob ? bearing(src[0], src[1], rcv[0], rcv[1])                       // Global: an azimuth; Cartesian: from the x axis
                          : 180.0 / 3.14159 * std::atan2(rcv[1] - src[1], rcv[0] - src[0]);
        int iterations = 0;
        e.theta_est = theta_max;
        double r, r_prev, d_theta = o.d_theta_big, d_phi = 10.0;
        bool theta_max_reached = false;
        while (std::fabs(d_phi) > o.azimuth_err_lim && iterations < 5) {
            r = r_rcvr; r_prev = r_rcvr;
            const double phi_rad = glob ? (90.0 - phi) * Pi / 180.0 : phi * Pi / 180.0;      // GeoAc_phi
            // Cartesian: theta <= theta_max (Eigenray.cpp:58); Global: theta < theta_max (Eigenray.Global.cpp:74)
            for (double theta = theta_min; glob ? theta < theta_max : theta <= theta_max; theta += d_theta) {
                if (theta + d_theta >= theta_max) theta_max_reached = true;
                bool brk;
                const double* s = lookup(theta * Pi / 180.0, phi_rad, 0, bounces, &brk);
                if (!s) {
                    if (iterations < 3) {          // fixed step: the rest of this fan is known now
                        int guard = 0;
                        for (double t = theta + d_theta; (glob ? t < theta_max : t <= theta_max) && guard < 100000; t += d_theta, guard++) { bool b2; lookup(t * Pi / 180.0, phi_rad, 0, bounces, &b2); }
                    }
                    return e;                      // incomplete
                }
                if (brk) { r = r_rcvr; r_prev = r_rcvr; }
                else if (glob) r = gc_distance(src[0], src[1], s[1] * 180.0 / Pi, s[2] * 180.0 / Pi);
                else r = std::sqrt(std::pow(s[0] - src[0], 2) + std::pow(s[1] - src[1], 2));
                if ((r - r_rcvr) * (r_prev - r_rcvr) < 0.0) {
                    if (iterations == 0) { e.theta_next = theta; e.next_known = true; }
                    if (glob) {
                        d_phi = bearing(src[0], src[1], rcv[0], rcv[1]);
                        d_phi -= bearing(src[0], src[1], s[1] * 180.0 / Pi, s[2] * 180.0 / Pi);
                    } else d_phi = (std::atan2(rcv[1] - src[1], rcv[0] - src[0]) - std::atan2(s[1] - src[1], s[0] - src[0])) * 180.0 / Pi;
                    while (d_phi > 180.0) d_phi -= 360.0;
                    while (d_phi < -180.0) d_phi += 360.0;
                    if (std::fabs(d_phi) < o.azimuth_err_lim) {
                        e.theta_est = theta - d_theta; e.phi_est = glob ? 90.0 - phi : phi; e.ok = true; e.complete = true;
                        return e;
                    }
                    phi += d_phi * 0.9;
                    theta_min = std::max(theta - 7.5, theta_min);
                    break;
                }
                if (iterations >= 3) d_theta = modify_d_theta(r - r_rcvr, (r - r_prev) / (2.0 * d_theta), o.d_theta_big, o.d_theta_small, glob ? 0.5 : 2.0);
                r_prev = r;
            }
            if (theta_max_reached) { e.theta_next = theta_max; e.next_known = true; break; }
            iterations++;
            if (iterations >= 1 && iterations < 3) d_theta = o.d_theta_big / 2.0;
        }
        e.complete = true;
        return e;
    }

    // ---- GeoAc_3DEigenray_LM, Eigenray.cpp:123-335 (search part; the attributes of a found eigenray come from one more trace) ----
    struct Lm { bool complete = false, found = false; double theta = 0, phi = 0; int iters = 0; };
    Lm lm(const double rcv[2], double theta, double phi, int bnc_cnt) {
        Lm out; out.theta = theta; out.phi = phi;
        double dr, dr_prev = 10000.0;
        const double tolerance = o.tolerance, theta_lim_step = 0.2, phi_lim_step = 0.2;
        double step_scalar = 1.0;
        long double x, y, dx, dy, dx_dt, dy_dt, dx_dp, dy_dp, det, dt = 0, dp = 0;
        double nu0_xy[2] = { 0, 0 };
        for (int n = 0; n <= o.iterations; n++) {
            out.iters = n;
            if (n == o.iterations) break;
            const double th = theta * Pi / 180.0, ph = phi * Pi / 180.0;
            if (strat && !glob) {
                const double nu0[3] = { std::cos(th) * std::cos(ph), std::cos(th) * std::sin(ph), std::sin(th) };
                const double M = 1.0 + (nu0[0] * Mx + nu0[1] * My + nu0[2] * 0.0);
                nu0_xy[0] = nu0[0] / M; nu0_xy[1] = nu0[1] / M;
            }
            bool brk;
            const double* s = lookup(th, ph, 1, bnc_cnt, &brk);
            if (!s) { out.theta = theta; out.phi = phi; return out; }          // incomplete
            if (brk) break;
            if (glob) {                                                        // Eigenray.Global.cpp:183-184 (x = lat, y = lon)
                x = s[1]; y = s[2];
                dr = gc_distance((double)(x * 180.0 / Pi), (double)(y * 180.0 / Pi), rcv[0], rcv[1]);
            } else {
                x = s[0]; dx = rcv[0] - x;
                y = s[1]; dy = rcv[1] - y;
                dr = (double)std::sqrt(dx * dx + dy * dy);
            }
            if (dr < tolerance) { out.found = true; break; }
            else if (n > 0 && dr > dr_prev) {
                theta -= dt * step_scalar;
                phi -= dp * step_scalar;
                step_scalar /= 2.0;
                if (std::sqrt(dt * dt + dp * dp) * step_scalar < 1.0e-12) break;
            } else {
                step_scalar = std::min(1.0, step_scalar * 1.25);
                if (glob) {                                                    // Eigenray.Global.cpp:283-294; dx, dy = d_lat, d_lon
                    const double rg = 6370.0 + z_grnd;
                    dx = rcv[0] * Pi / 180.0 - x;
                    dy = rcv[1] * Pi / 180.0 - y;
                    dx_dt = s[7] - 1.0 / rg * s[4] / s[3] * s[6];
                    dx_dp = s[13] - 1.0 / rg * s[4] / s[3] * s[12];
                    dy_dt = s[8] - 1.0 / (rg * std::cos(x)) * s[5] / s[3] * s[6];
                    dy_dp = s[14] - 1.0 / (rg * std::cos(x)) * s[5] / s[3] * s[12];
                    det = dx_dt * dy_dp - dx_dp * dy_dt;
                    dt = (dy_dp * dx - dx_dp * dy) / det * 180.0 / Pi;
                    dp = (-dy_dt * dx + dx_dt * dy) / det * 180.0 / Pi;
                } else if (strat) {
                    dx_dt = s[4] - nu0_xy[0] / s[3] * s[6];
                    dy_dt = s[5] - nu0_xy[1] / s[3] * s[6];
                    dx_dp = s[8] - nu0_xy[0] / s[3] * s[10];
                    dy_dp = s[9] - nu0_xy[1] / s[3] * s[10];
                } else {
                    dx_dt = s[6] - s[3] / s[5] * s[8];
                    dy_dt = s[7] - s[4] / s[5] * s[8];
                    dx_dp = s[12] - s[3] / s[5] * s[14];
                    dy_dp = s[13] - s[4] / s[5] * s[14];
                }
                if (!glob) {
                    det = dx_dt * dy_dp - dx_dp * dy_dt;
                    dt = 1.0 / det * (dy_dp * dx - dx_dp * dy) * 180.0 / Pi;
                    dp = 1.0 / det * (dx_dt * dy - dy_dt * dx) * 180.0 / Pi;
                }
                if (dt > theta_lim_step) dt = theta_lim_step;
                if (dp > phi_lim_step) dp = phi_lim_step;
                if (dt < -theta_lim_step) dt = -theta_lim_step;
                if (dp < -phi_lim_step) dp = -phi_lim_step;
                theta += dt * step_scalar;
                phi += dp * step_scalar;
                dr_prev = dr;
            }
        }
        out.complete = true; out.theta = theta; out.phi = phi;
        return out;
    }

    // one GeoAc_EstimateEigenray call (+ its LM) of the chain of (receiver, bounce count)
    struct Call { int rcvr, n_bnc; Est e; Lm l; };

    // Replays every chain against the cache.  Returns true when nothing is missing.
    // direct != nullptr: -eig_direct (GeoAc3D_RunEigDirect, GeoAc3D_main.cpp:546-601): no estimate, one LM search per entry
    // from the caller's { theta_est, phi_est [deg from the x axis], bounces }.
    bool replay(int n_rcvr, const double* rcv_xy, const double* direct, std::vector<Call>& calls) {
        calls.clear();
        bool all = true;
        if (direct) {
            for (int ir = 0; ir < n_rcvr; ir++) {
                const double rcv[2] = { rcv_xy[2 * ir], rcv_xy[2 * ir + 1] };
                Call c; c.rcvr = ir; c.n_bnc = (int)direct[3 * ir + 2];
                c.e.complete = c.e.ok = c.e.next_known = true; c.e.theta_est = direct[3 * ir]; c.e.phi_est = direct[3 * ir + 1]; c.e.theta_next = 0.0;
                c.l = lm(rcv, c.e.theta_est, c.e.phi_est, c.n_bnc);
                all = all && c.l.complete;
                calls.push_back(c);
            }
            return all;
        }
        for (int ir = 0; ir < n_rcvr; ir++) {
            const double rcv[2] = { rcv_xy[2 * ir], rcv_xy[2 * ir + 1] };
            for (int n_bnc = o.bnc_min; n_bnc <= o.bnc_max; n_bnc++) {
                double theta_start = o.theta_min;
                int guard = 0;
                while (theta_start < o.theta_max && guard++ < 100000) {
                    Call c; c.rcvr = ir; c.n_bnc = n_bnc;
                    c.e = estimate(rcv, theta_start, o.theta_max, n_bnc);
                    if (c.e.complete && c.e.ok) c.l = lm(rcv, c.e.theta_est, c.e.phi_est, n_bnc);
                    const bool done = c.e.complete && (!c.e.ok || c.l.complete);
                    all = all && done;
                    calls.push_back(c);
                    if (!c.e.next_known) break;                   // cannot place the next call yet (or theta_next was never set, as in the reference)
                    theta_start = c.e.theta_next;
                }
            }
        }
        return all;
    }

    int trace_requests(std::string& err) {
        for (int amp = 0; amp < 2; amp++) {
            std::vector<double> th, ph; std::vector<RayKey> keys; int nb = 0;
            for (auto& kv : want) if (kv.first.amp == amp) {
                th.push_back(kv.second.theta); ph.push_back(kv.second.phi);
                keys.push_back(kv.first); nb = std::max(nb, kv.second.nb);
            }
            if (th.empty()) continue;
            geoac_params p; int rc = geoac_get_params(ctx, &p); if (rc) return rc;
            p.bounces = nb; p.calc_amp = amp; p.accum_per_segment = 0;
            rc = geoac_set_params(ctx, &p); if (rc) return rc;
            const int64_t n = (int64_t)th.size(), n_rec = nb + 1, slots = n * n_rec;
            std::vector<double> rec((size_t)GEOAC_NFIELDS * slots);
            std::vector<int32_t> st((size_t)slots), ns((size_t)slots);
            rc = geoac_trace(ctx, n, th.data(), ph.data(), rec.data(), st.data(), ns.data());
            if (rc) { err = geoac_last_error(ctx); return rc; }
            rays_traced += n;
            for (int64_t i = 0; i < n; i++) {
                RayVal v; v.nb = nb; v.status.assign(st.begin() + i * n_rec, st.begin() + (i + 1) * n_rec);
                v.state.assign((size_t)n_rec * 18, 0.0);
                for (int b = 0; b < n_rec; b++) for (int f = 0; f < 18; f++) v.state[(size_t)b * 18 + f] = rec[(size_t)f * slots + i * n_rec + b];
                cache[keys[i]] = std::move(v);
            }
        }
        want.clear();
        return GEOAC_OK;
    }
};

}  // namespace

extern "C" int geoac_default_eig_opts(geoac_eig_opts* o) {
    if (!o) return GEOAC_ERR_BAD_ARG;
    std::memset(o, 0, sizeof *o);
    o->theta_min = 0.5; o->theta_max = 45.0; o->bnc_min = 0; o->bnc_max = 0; o->iterations = 25;      // GeoAc3D_main.cpp:461-464
    o->azimuth_err_lim = 2.0; o->d_theta_big = 0.25; o->d_theta_small = 0.002; o->tolerance = 0.1;   // Eigenray.cpp:20-21,139
    o->max_rounds = 4096; o->src_lat_deg = 30.0; o->src_lon_deg = 0.0;                              // GeoAcGlobal_main.cpp:497
    return GEOAC_OK;
}

static int run_search(geoac_ctx* ctx, const geoac_eig_opts* opts, int n_rcvr, const double* rcvr_xy, const double* direct,
                      int64_t cap_rows, double* rows, int64_t* n_rows, int64_t* stats) {
    if (!ctx || !opts || !rcvr_xy || !n_rows || n_rcvr < 0 || (cap_rows > 0 && !rows)) return GEOAC_ERR_BAD_ARG;
    const int variant = geoac_get_variant(ctx);
    if (variant == GEOAC_2D) return GEOAC_ERR_BAD_ARG;                                    // GeoAc2D has no eigenray search
    if (opts->bnc_min < 0 || opts->bnc_max < opts->bnc_min || opts->iterations < 0 || !(opts->d_theta_big > 0.0)) return GEOAC_ERR_BAD_ARG;
    geoac_params user; int rc = geoac_get_params(ctx, &user); if (rc) return rc;
    Search S; S.ctx = ctx; S.variant = variant; S.strat = variant == GEOAC_3D; S.o = *opts;
    S.glob = variant == GEOAC_GLOBAL || variant == GEOAC_GLOBAL_RNGDEP; S.z_grnd = user.z_grnd;
    if (S.glob) {
        // Source_Loc = (lat, lon) in degrees as the Global mains hold it (GeoAcGlobal_main.cpp:497); the rays start from
        // lat*Pi/180, lon*Pi/180 (Eigenray.Global.cpp:78), which is written into the context for the duration of the search
        S.src[0] = opts->src_lat_deg; S.src[1] = opts->src_lon_deg; S.src[2] = std::max(user.z_grnd, user.src[0]);
        geoac_params p = user; p.src[0] = S.src[2]; p.src[1] = S.src[0] * Pi / 180.0; p.src[2] = S.src[1] * Pi / 180.0;
        rc = geoac_set_params(ctx, &p); if (rc) return rc;
    } else { S.src[0] = user.src[0]; S.src[1] = user.src[1]; S.src[2] = std::max(user.z_grnd, user.src[2]); }
    if (S.strat) {
        double a[4]; rc = geoac_source_state(ctx, a); if (rc) return rc;
        S.Mx = a[1] / a[0]; S.My = a[2] / a[0];
    }
    std::vector<Search::Call> calls;
    std::string err;
    const int max_rounds = opts->max_rounds > 0 ? opts->max_rounds : 4096;
    bool done = false;
    for (S.rounds = 0; S.rounds < max_rounds; S.rounds++) {
        done = S.replay(n_rcvr, rcvr_xy, direct, calls);
        if (done) break;
        rc = S.trace_requests(err);
        if (rc) { geoac_set_params(ctx, &user); return rc; }
    }
    if (!done) { geoac_set_params(ctx, &user); return GEOAC_ERR_TOO_LARGE; }
    // attributes of the eigenrays found: the per-segment sums of the reference's final pass (Eigenray.cpp:203-244), one batch
    std::vector<int> found;
    for (size_t i = 0; i < calls.size(); i++) if (calls[i].e.ok && calls[i].l.found) found.push_back((int)i);
    std::vector<double> rec; std::vector<int32_t> st; int nbmax = 0; int64_t slots = 0;
    if (!found.empty()) {
        std::vector<double> th, ph;
        for (int i : found) { th.push_back(calls[i].l.theta * Pi / 180.0); ph.push_back(calls[i].l.phi * Pi / 180.0); nbmax = std::max(nbmax, calls[i].n_bnc); }
        geoac_params p; rc = geoac_get_params(ctx, &p); if (rc) { geoac_set_params(ctx, &user); return rc; }      // keeps the search's source
        p.bounces = nbmax; p.calc_amp = 1; p.accum_per_segment = 1;
        rc = geoac_set_params(ctx, &p);
        slots = (int64_t)found.size() * (nbmax + 1);
        rec.resize((size_t)GEOAC_NFIELDS * slots); st.resize((size_t)slots); std::vector<int32_t> ns((size_t)slots);
        if (!rc) rc = geoac_trace(ctx, (int64_t)found.size(), th.data(), ph.data(), rec.data(), st.data(), ns.data());
        if (rc) { geoac_set_params(ctx, &user); return rc; }
        S.rays_traced += (int64_t)found.size();
    }
    geoac_set_params(ctx, &user);
    *n_rows = (int64_t)calls.size();
    int64_t fi = 0;
    for (size_t i = 0; i < calls.size(); i++) {
        const Search::Call& c = calls[i];
        const bool is_found = c.e.ok && c.l.found;
        if ((int64_t)i < cap_rows) {
            double* r = rows + i * GEOAC_EIG_NF;
            for (int k = 0; k < GEOAC_EIG_NF; k++) r[k] = 0.0;
            r[0] = c.rcvr; r[1] = c.n_bnc; r[2] = c.e.ok ? 1.0 : 0.0; r[3] = c.e.theta_est; r[4] = c.e.ok ? c.e.phi_est : 0.0; r[5] = c.e.theta_next;
            r[6] = is_found ? 1.0 : 0.0;
            if (c.e.ok) { r[7] = c.l.theta; r[8] = c.l.phi; r[16] = c.l.iters; }
            if (is_found) {
                const int64_t slot = fi * (nbmax + 1) + c.n_bnc;
                auto F = [&](int f) { return rec[(size_t)f * slots + slot]; };
                const double rx = rcvr_xy[2 * c.rcvr], ry = rcvr_xy[2 * c.rcvr + 1];
                const double tt = F(GEOAC_F_TRAVELTIME);
                r[9] = tt;
                r[11] = F(GEOAC_F_AMPLITUDE); r[12] = F(GEOAC_F_ATTEN); r[13] = F(GEOAC_F_INCLINATION);
                double back_az, dev;
                if (S.glob) {                                                                                      // Eigenray.Global.cpp:238-241,266-273
                    r[10] = Search::gc_distance(S.src[0], S.src[1], rx, ry) / tt;
                    if (variant == GEOAC_GLOBAL_RNGDEP) r[13] = -r[13];          // the record follows that main's +asin (App. A-16); the search prints -asin
                    back_az = 90.0 - std::atan2(-F(4), -F(5)) * 180.0 / Pi;
                    dev = back_az - Search::bearing(rx, ry, S.src[0], S.src[1]);
                    if (dev > 180.0) dev -= 360.0;
                    if (dev < -180.0) dev += 360.0;
                } else {
                    r[10] = std::sqrt(std::pow(F(0) - S.src[0], 2) + std::pow(F(1) - S.src[1], 2)) / tt;           // celerity, Eigenray.cpp:260
                    back_az = S.strat ? (90.0 - (c.l.phi * Pi / 180.0) * 180.0 / Pi) + 180.0                        // :239-244
                                      : 90.0 - std::atan2(-F(4), -F(3)) * 180.0 / Pi;
                    dev = back_az - (90.0 - std::atan2(S.src[1] - ry, S.src[0] - rx) * 180.0 / Pi);
                    while (back_az > 180.0) back_az -= 360.0;
                    while (back_az < -180.0) back_az += 360.0;
                    while (dev > 180.0) dev -= 360.0;
                    while (dev < -180.0) dev += 360.0;
                }
                r[14] = back_az; r[15] = dev;
                r[17] = st[(size_t)slot];
            }
        }
        if (is_found) fi++;
    }
    if (stats) { stats[0] = S.rounds; stats[1] = S.rays_traced; stats[2] = (int64_t)found.size(); }
    return (int64_t)calls.size() <= cap_rows ? GEOAC_OK : GEOAC_ERR_TOO_LARGE;
}

extern "C" int geoac_eigenray_search(geoac_ctx* ctx, const geoac_eig_opts* opts, int n_rcvr, const double* rcvr_xy,
                                     int64_t cap_rows, double* rows, int64_t* n_rows, int64_t* stats) {
    return run_search(ctx, opts, n_rcvr, rcvr_xy, nullptr, cap_rows, rows, n_rows, stats);
}

extern "C" int geoac_eigenray_direct(geoac_ctx* ctx, const geoac_eig_opts* opts, int n, const double* rcvr_xy, const double* estimates,
                                     double* rows, int64_t* stats) {
    if (!estimates) return GEOAC_ERR_BAD_ARG;
    for (int i = 0; i < n; i++) if (!(estimates[3 * i + 2] >= 0.0) || estimates[3 * i + 2] > 1000.0) return GEOAC_ERR_BAD_ARG;
    int64_t n_rows = 0;
    return run_search(ctx, opts, n, rcvr_xy, estimates, n, rows, &n_rows, stats);
}
