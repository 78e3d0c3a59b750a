// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, not product code.
//
// A tiny `main` that is linked against the UNMODIFIED reference sources where they lie under
// /root/reference/Code (see oracle/Makefile; outputs go to oracle/_ref/, git-ignored).  It repeats the call
// sequence of the reference's `-prop` loops (Code/GeoAc3D_main.cpp:226-304, Code/GeoAc2D_main.cpp:170-229,
// Code/GeoAcGlobal_main.cpp:241-322, Code/GeoAc3D.RngDep_main.cpp:244-323, Code/GeoAcGlobal.RngDep_main.cpp:251-331)
// but, instead of printing 6-digit text, dumps one raw-double record per (ray, bounce) so that parity can be
// checked to full FP64 precision.  One variant per executable (the reference defines the same symbols once per
// variant), selected with -DREF_2D / -DREF_3D / -DREF_GLOBAL / -DREF_3DRNGDEP / -DREF_GLOBALRNGDEP.
//
// Usage:  ref_<variant> <out.bin> <profile args...> [key=value ...]
//   profile args: 1 file (stratified) or 3 (prefix, loc file 1, loc file 2) for the range-dependent variants.
//   keys: theta_min theta_max theta_step phi_min phi_max phi_step azimuth bounces z_src x_src y_src lat_src lon_src
//         freq abs_coeff z_grnd CalcAmp alt_max rng_max accum_mode (1 = per-segment post pass, i.e. WriteRays=True;
//         0 = GeoAc_TravelTime(k), i.e. WriteRays=False) stride offset (ray subsampling) profile_format quiet
//
// Record layout (REC_NF doubles) mirrors include/geoac_b200.h field ids.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <chrono>
#include <vector>
#include <string>
#include <algorithm>
#include <sys/resource.h>
#include <unistd.h>

#include "GeoAc/GeoAc.Parameters.h"
#include "Atmo/Atmo_State.h"
#include "GeoAc/GeoAc.EquationSets.h"
#include "GeoAc/GeoAc.Solver.h"
#include "GeoAc/GeoAc.Interface.h"

#if defined(REF_3DRNGDEP) || defined(REF_GLOBALRNGDEP)
#define REF_RNGDEP 1
void Spline_Multi_G2S(char*, char*, char*, char*);
#else
void Spline_Single_G2S(char*, char*);
#endif

enum { F_TT = 18, F_ATT = 19, F_TURN = 20, F_AMP = 21, F_INCL = 22, F_BACKAZ = 23, F_AUX = 24, F_MARGIN = 25,
       F_STATUS = 26, F_NSTEPS = 27, REC_NF = 32 };

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv) {
    // the range-dependent slope builders put two full-grid VLAs on the stack (SURVEY 8c caveat 1)
    struct rlimit rl; getrlimit(RLIMIT_STACK, &rl);
    if (rl.rlim_cur != RLIM_INFINITY && rl.rlim_cur < (rlim_t)1 << 33) {
        rl.rlim_cur = rl.rlim_max; setrlimit(RLIMIT_STACK, &rl);
        if (!getenv("REF_DRIVER_REEXEC")) { setenv("REF_DRIVER_REEXEC", "1", 1); execv("/proc/self/exe", argv); }
    }
#ifdef REF_RNGDEP
    const int nprof = 3;
#else
    const int nprof = 1;
#endif
    if (argc < 2 + nprof) { fprintf(stderr, "usage: %s out.bin profile... key=value...\n", argv[0]); return 2; }
    const char* out_path = argv[1];
    double theta_min = 0.5, theta_max = 45.0, theta_step = 0.5;
    double phi_min = -90.0, phi_max = -90.0, phi_step = 1.0;
    int bounces = 2, accum_mode = 0, stride = 1, offset = 0, quiet = 1, path_stride = 0, caustics = 0;
    double x_src = 0.0, y_src = 0.0, z_src = 0.0, lat_src = 30.0, lon_src = 0.0;
    bool have_src = false;
    bool CalcAmp = true;
    double freq = 0.1;
    char* fmt = (char*)"zTuvdp";
    z_grnd = 0.0; tweak_abs = 0.3;
    int first_kv = 2 + nprof;
    for (int i = first_kv; i < argc; i++) if (!strncmp(argv[i], "profile_format=", 15)) fmt = argv[i] + 15;

#ifndef REF_RNGDEP
    Spline_Single_G2S(argv[2], fmt);            // as in the stratified mains: load BEFORE parsing (z_grnd = 0 taper)
#endif
    for (int i = first_kv; i < argc; i++) {
        char* a = argv[i];
        if      (!strncmp(a, "theta_min=", 10))  theta_min = atof(a + 10);
        else if (!strncmp(a, "theta_max=", 10))  theta_max = atof(a + 10);
        else if (!strncmp(a, "theta_step=", 11)) theta_step = atof(a + 11);
        else if (!strncmp(a, "phi_min=", 8))     phi_min = atof(a + 8);
        else if (!strncmp(a, "phi_max=", 8))     phi_max = atof(a + 8);
        else if (!strncmp(a, "phi_step=", 9))    phi_step = atof(a + 9);
        else if (!strncmp(a, "azimuth=", 8))     { phi_min = phi_max = atof(a + 8); phi_step = 1.0; }
        else if (!strncmp(a, "bounces=", 8))     bounces = atoi(a + 8);
        else if (!strncmp(a, "x_src=", 6))       x_src = atof(a + 6);
        else if (!strncmp(a, "y_src=", 6))       y_src = atof(a + 6);
        else if (!strncmp(a, "z_src=", 6))       z_src = atof(a + 6);
        else if (!strncmp(a, "lat_src=", 8))     { lat_src = atof(a + 8); have_src = true; }
        else if (!strncmp(a, "lon_src=", 8))     { lon_src = atof(a + 8); have_src = true; }
        else if (!strncmp(a, "freq=", 5))        freq = atof(a + 5);
        else if (!strncmp(a, "abs_coeff=", 10))  tweak_abs = std::max(0.0, atof(a + 10));
        else if (!strncmp(a, "z_grnd=", 7))      z_grnd = atof(a + 7);
        else if (!strncmp(a, "CalcAmp=", 8))     CalcAmp = atoi(a + 8) != 0;
        else if (!strncmp(a, "alt_max=", 8))     GeoAc_vert_limit = atof(a + 8);
#ifndef REF_RNGDEP
        else if (!strncmp(a, "rng_max=", 8))     GeoAc_range_limit = atof(a + 8);
#endif
        else if (!strncmp(a, "accum_mode=", 11)) accum_mode = atoi(a + 11);
        else if (!strncmp(a, "caustics=", 9))    caustics = atoi(a + 9);           // WriteCaustics=True: rows where the Jacobian changes sign
        else if (!strncmp(a, "path_stride=", 12)) path_stride = atoi(a + 12);   // WriteRays=True: one raypath row every N steps (the mains use 25)
        else if (!strncmp(a, "stride=", 7))      stride = atoi(a + 7);
        else if (!strncmp(a, "offset=", 7))      offset = atoi(a + 7);
        else if (!strncmp(a, "quiet=", 6))       quiet = atoi(a + 6);
        else if (!strncmp(a, "profile_format=", 15)) {}
        else { fprintf(stderr, "unknown key %s\n", a); return 2; }
    }
    z_src = std::max(z_grnd, z_src);
#ifdef REF_RNGDEP
    // range-dependent mains: grid is loaded AFTER parsing and SetPropRegion then overwrites the CLI limits
    double t_load0 = now_s();
    Spline_Multi_G2S(argv[2], argv[3], argv[4], fmt);
    GeoAc_SetPropRegion();
    double t_load = now_s() - t_load0;
#else
    double t_load = 0.0;
#endif
#ifdef REF_GLOBALRNGDEP
    if (!have_src) {   // Code/GeoAcGlobal.RngDep_main.cpp:135-137 default source = grid midpoint (degrees)
        lat_src = (GeoAc_lat_min_limit + GeoAc_lat_max_limit) / 2.0 * 180.0 / Pi;
        lon_src = (GeoAc_lon_min_limit + GeoAc_lon_max_limit) / 2.0 * 180.0 / Pi;
    }
#endif
    (void)have_src; (void)x_src; (void)y_src; (void)lat_src; (void)lon_src;
#ifdef REF_2D
    accum_mode = 1;                             // the 2D main always uses the per-segment post pass
#endif
    GeoAc_ConfigureCalcAmp(CalcAmp);

    int length = GeoAc_ray_limit * int(1.0 / (GeoAc_ds_min * 10));   // == RK4 step_limit (avoids Appendix A-12 overrun)
    double** solution; GeoAc_BuildSolutionArray(solution, length + 2);   // +2: RK4 returns `length` at the step limit and callers read solution[k]

    std::vector<double> angles;                 // theta_deg, phi_deg per traced ray
    std::vector<double> recs;
    std::vector<double> caus;                   // caustic rows: ray, state[0..2], travel time, bounce, step
    std::vector<double> paths;                  // raypath rows: ray, state[0..2], amplitude, attenuation, travel time, bounce, step
    long total_steps = 0; double t_rk4 = 0.0, t_post = 0.0, t_all0 = now_s();
    long ray_index = 0, n_traced = 0;
#ifdef REF_2D
    phi_step = 1.0;                              // 2D has a single azimuth (phi_min == phi_max == azimuth)
#endif
    for (double phi = phi_min; phi <= phi_max; phi += phi_step) {
    for (double theta = theta_min; theta <= theta_max; theta += theta_step) {
        long idx = ray_index++;
        if (idx % stride != offset) continue;
        n_traced++;
        angles.push_back(theta); angles.push_back(phi);
        size_t base = recs.size(); recs.resize(base + (size_t)(bounces + 1) * REC_NF, 0.0);

        GeoAc_theta = theta * Pi / 180.0;
        GeoAc_phi = Pi / 2.0 - phi * Pi / 180.0;
#if defined(REF_2D)
        GeoAc_SetInitialConditions(solution, 0.0, z_src);
#elif defined(REF_3D) || defined(REF_3DRNGDEP)
        GeoAc_SetInitialConditions(solution, x_src, y_src, z_src);
#else
        GeoAc_SetInitialConditions(solution, z_src, lat_src * Pi / 180.0, lon_src * Pi / 180.0);
#endif
        double travel_time_sum = 0.0, attenuation = 0.0, z_max = 0.0;
        if (!quiet) printf("ray theta=%g phi=%g\n", theta, phi);
        for (int bnc = 0; bnc <= bounces; bnc++) {
            bool BreakCheck; int k;
            double t0 = now_s();
            k = GeoAc_Propagate_RK4(solution, BreakCheck);
            double t1 = now_s();
            if (accum_mode) {
                double D = 0.0, D_prev = 0.0;
                if (caustics) D_prev = GeoAc_Jacobian(solution, 1);                 // Code/GeoAc3D_main.cpp:245
                for (int m = 1; m < k; m++) {
                    if (caustics) D = GeoAc_Jacobian(solution, m);
                    GeoAc_TravelTimeSegment(travel_time_sum, solution, m - 1, m);
                    GeoAc_SB_AttenSegment(attenuation, solution, m - 1, m, freq);
                    if (path_stride > 0 && m % path_stride == 0) {            // Code/GeoAc3D_main.cpp:254-262 (and the sibling mains)
                        const double row[9] = { (double)(n_traced - 1), solution[m][0], solution[m][1], solution[m][2],
                                                CalcAmp ? GeoAc_Amplitude(solution, m) : 0.0, attenuation, travel_time_sum, (double)bnc, (double)m };
                        paths.insert(paths.end(), row, row + 9);
                    }
                    if (caustics && D * D_prev < 0.0) {                              // :263-268
                        const double row[7] = { (double)(n_traced - 1), solution[m][0], solution[m][1], solution[m][2], travel_time_sum, (double)bnc, (double)m };
                        caus.insert(caus.end(), row, row + 7);
                    }
                    if (caustics) D_prev = D;
                }
            } else {
                travel_time_sum += GeoAc_TravelTime(solution, k);
                attenuation += GeoAc_SB_Atten(solution, k, freq);
            }
            double t2 = now_s();
            t_rk4 += t1 - t0; t_post += t2 - t1; total_steps += k;
            double* rec = &recs[base + (size_t)bnc * REC_NF];
            rec[F_NSTEPS] = k;
            if (BreakCheck) { rec[F_STATUS] = 2.0; break; }
            if (k >= length) { rec[F_STATUS] = 3.0; break; }   // step limit reached (no ground hit, no break)
            rec[F_STATUS] = 1.0;
#if defined(REF_3DRNGDEP) || defined(REF_GLOBALRNGDEP)
            z_max = 0.0;                          // Appendix A-3: reset per bounce in the range-dependent mains
#endif
#if defined(REF_2D)
            for (int m = 0; m < k; m++) z_max = std::max(z_max, solution[m][1]);
#elif defined(REF_3D) || defined(REF_3DRNGDEP)
            for (int m = 0; m < k; m++) z_max = std::max(z_max, solution[m][2]);
#else
            for (int m = 0; m < k; m++) z_max = std::max(z_max, solution[m][0] - r_earth);
#endif
            for (int i = 0; i < GeoAc_EqCnt; i++) rec[i] = solution[k][i];
            rec[F_TT] = travel_time_sum; rec[F_ATT] = attenuation; rec[F_TURN] = z_max;
            rec[F_AMP] = CalcAmp ? GeoAc_Amplitude(solution, k) : 0.0;
#if defined(REF_2D)
            rec[F_INCL] = -theta; rec[F_BACKAZ] = 0.0;
            rec[F_MARGIN] = (solution[k][1] - z_grnd) / fabs(solution[k][1] - solution[k - 1][1]);
#elif defined(REF_3D)
            {
                double back_az = phi + 180.0;
                rec[F_INCL] = -asin(c(solution[k][0], solution[k][1], z_grnd) / c(x_src, y_src, z_src) * solution[k][3]) * 180.0 / Pi;
                while (back_az > 180.0) back_az -= 360.0;
                while (back_az < -180.0) back_az += 360.0;
                rec[F_BACKAZ] = back_az;
                rec[F_MARGIN] = (solution[k][2] - z_grnd) / fabs(solution[k][2] - solution[k - 1][2]);
            }
#elif defined(REF_3DRNGDEP)
            {
                rec[F_INCL] = -asin(c(solution[k][0], solution[k][1], z_grnd) / c(x_src, y_src, z_src) * solution[k][5]) * 180.0 / Pi;
                double back_az = 90.0 - atan2(-solution[k][4], -solution[k][3]) * 180.0 / Pi;
                while (back_az < -180.0) back_az += 360.0;
                while (back_az > 180.0) back_az -= 360.0;
                rec[F_BACKAZ] = back_az;
                rec[F_MARGIN] = (solution[k][2] - z_grnd) / fabs(solution[k][2] - solution[k - 1][2]);
            }
#else
            {
                double GC_Dist1 = pow(sin((solution[k][1] - lat_src * Pi / 180.0) / 2.0), 2);
                double GC_Dist2 = cos(lat_src * Pi / 180.0) * cos(solution[k][1]) * pow(sin((solution[k][2] - lon_src * Pi / 180.0) / 2.0), 2);
                double incl = asin(c(solution[k][0], solution[k][1], solution[k][2]) / c(r_earth + z_src, lat_src * Pi / 180.0, lon_src * Pi / 180.0) * solution[k][3]) * 180.0 / Pi;
#if defined(REF_GLOBAL)
                incl = -incl;                     // Appendix A-16: sign differs between the two Global mains
#endif
                double back_az = 90.0 - atan2(-solution[k][4], -solution[k][5]) * 180.0 / Pi;
                if (back_az < -180.0) back_az += 360.0;
                if (back_az > 180.0) back_az -= 360.0;
                rec[F_INCL] = incl; rec[F_BACKAZ] = back_az;
                rec[F_AUX] = 2.0 * r_earth * asin(sqrt(GC_Dist1 + GC_Dist2)) / travel_time_sum;   // celerity
                rec[F_MARGIN] = (solution[k][0] - (r_earth + z_grnd)) / fabs(solution[k][0] - solution[k - 1][0]);
            }
#endif
            GeoAc_SetReflectionConditions(solution, k);
        }
    }
    }
    double t_all = now_s() - t_all0;

    FILE* f = fopen(out_path, "wb");
    if (!f) { perror("open out"); return 1; }
    double hdr[8] = { 20251018.0, (double)n_traced, (double)(bounces + 1), (double)REC_NF, (double)GeoAc_EqCnt,
                      (double)total_steps, t_rk4, t_post };
    fwrite(hdr, sizeof(double), 8, f);
    fwrite(angles.data(), sizeof(double), angles.size(), f);
    fwrite(recs.data(), sizeof(double), recs.size(), f);
    fclose(f);
    if (caustics) {
        std::string cp = std::string(out_path) + ".caus";
        FILE* g = fopen(cp.c_str(), "wb");
        if (!g) { perror("open caustics out"); return 1; }
        fwrite(caus.data(), sizeof(double), caus.size(), g);
        fclose(g);
    }
    if (path_stride > 0) {
        std::string pp = std::string(out_path) + ".path";
        FILE* g = fopen(pp.c_str(), "wb");
        if (!g) { perror("open path out"); return 1; }
        fwrite(paths.data(), sizeof(double), paths.size(), g);
        fclose(g);
    }
    printf("{\"rays\": %ld, \"steps\": %ld, \"t_rk4_s\": %.6f, \"t_post_s\": %.6f, \"t_trace_s\": %.6f, \"t_load_s\": %.3f, \"eq_cnt\": %d, "
           "\"vert_limit\": %.17g}\n",
           n_traced, total_steps, t_rk4, t_post, t_all, t_load, GeoAc_EqCnt, GeoAc_vert_limit);
    return 0;
}
