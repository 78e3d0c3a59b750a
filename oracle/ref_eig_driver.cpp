// oracle/ref_eig_driver.cpp -- TEST INFRASTRUCTURE, not product code.
//
// `main` linked against the UNMODIFIED reference eigenray search (Code/GeoAc/GeoAc.Eigenray.cpp) and the 3-D stratified or
// range-dependent Cartesian sources where they lie under /root/reference/Code (oracle/Makefile: ref_eig3d, ref_eig3drngdep).
// It repeats the loop of GeoAc3D_RunEigSearch (Code/GeoAc3D_main.cpp:531-541; RngDep main likewise) -- for every bounce
// count, GeoAc_EstimateEigenray from theta_start, then GeoAc_3DEigenray_LM on success, theta_start = theta_next -- and
// dumps one raw-double row per GeoAc_EstimateEigenray call:
//   { n_bnc, estimate_ok, theta_est, phi_est, theta_next, eigenray_found, theta_final, phi_final }      (angles in degrees,
//   phi measured from the x axis as inside the reference).  The reference's own text outputs (<title>_results.dat,
//   <title>_Eigenray-N.dat) are written into `workdir`.
//
// Usage: ref_eig3d | ref_eigglobal <out.bin> <workdir> <profile>  [key=value ...]            (stratified)
//        ref_eig3drngdep | ref_eigglobalrngdep <out.bin> <workdir> <prefix> <loc_1> <loc_2> [key=value ...]
//   keys: theta_min theta_max bnc_min bnc_max bounces x_src y_src z_src x_rcvr y_rcvr (Global: lat_src lon_src lat_rcvr lon_rcvr,
//         degrees; GeoAcGlobal_main.cpp:522-526) azimuth_err_lim iterations freq abs_coeff z_grnd alt_max rng_max profile_format verbose
// The Global variants link Code/GeoAc/GeoAc.Eigenray.Global.cpp instead; phi columns are then 90 - azimuth as that file holds them.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <string>
#include <algorithm>
#include <chrono>
#include <sys/resource.h>
#include <unistd.h>

#include "GeoAc/GeoAc.Parameters.h"
#include "Atmo/Atmo_State.h"
#include "GeoAc/GeoAc.EquationSets.h"
#include "GeoAc/GeoAc.Solver.h"
#include "GeoAc/GeoAc.Interface.h"
#include "GeoAc/GeoAc.Eigenray.h"

#if defined(REF_3DRNGDEP) || defined(REF_GLOBALRNGDEP)
#define REF_RNGDEP 1
void Spline_Multi_G2S(char*, char*, char*, char*);
#else
void Spline_Single_G2S(char*, char*);
#endif
#if defined(REF_GLOBAL) || defined(REF_GLOBALRNGDEP)
#define REF_GLOB 1
#endif

int main(int argc, char** argv) {
    struct rlimit rl; getrlimit(RLIMIT_STACK, &rl);             // Set_Slopes_Multi's stack VLAs (SURVEY 8c caveat 1)
    if (rl.rlim_cur != RLIM_INFINITY && rl.rlim_cur < (rlim_t)1 << 33) {
        rl.rlim_cur = rl.rlim_max; setrlimit(RLIMIT_STACK, &rl);
        if (!getenv("REF_DRIVER_REEXEC")) { setenv("REF_DRIVER_REEXEC", "1", 1); execv("/proc/self/exe", argv); }
    }
#ifdef REF_RNGDEP
    const int nprof = 3;
#else
    const int nprof = 1;
#endif
    if (argc < 3 + nprof) { fprintf(stderr, "usage: %s out.bin workdir profile... key=value...\n", argv[0]); return 2; }
    std::string out_path = argv[1];
    if (out_path[0] != '/') { char cwd[4096]; if (getcwd(cwd, sizeof cwd)) out_path = std::string(cwd) + "/" + out_path; }
    std::vector<std::string> prof;
    for (int i = 0; i < nprof; i++) {
        std::string p = argv[3 + i];
        if (p[0] != '/') { char cwd[4096]; if (getcwd(cwd, sizeof cwd)) p = std::string(cwd) + "/" + p; }
        prof.push_back(p);
    }
    if (chdir(argv[2]) != 0) { perror("chdir"); return 2; }

#ifdef REF_GLOB
    double Source_Loc[3] = { 30.0, 0.0, 0.0 }, Receiver_Loc[2] = { 30.0, -2.5 };      // (lat, lon) [deg], z: GeoAcGlobal_main.cpp:497-498
#else
    double Source_Loc[3] = { 0.0, 0.0, 0.0 }, Receiver_Loc[2] = { -250.0, 0.0 };
#endif
    double theta_min = 0.5, theta_max = 45.0, azimuth_err_lim = 2.0, freq = 0.1;
    int bnc_min = 0, bnc_max = 0, iterations = 25, direct = 0;
    double theta_direct = 0.5, phi_direct = 45.0;      // -eig_direct: theta_est= and phi_est= (phi_est as an azimuth, like the mains)
    bool have_phi = false;
    char* fmt = (char*)"zTuvdp";
    verbose_output = false; z_grnd = 0.0; tweak_abs = 0.3;
    const int first_kv = 3 + nprof;
    for (int i = first_kv; i < argc; i++) if (!strncmp(argv[i], "profile_format=", 15)) fmt = argv[i] + 15;
#ifndef REF_RNGDEP
    Spline_Single_G2S((char*)prof[0].c_str(), fmt);               // as in the main: load before parsing
#endif
    for (int i = first_kv; i < argc; i++) {
        char* a = argv[i];
        if      (!strncmp(a, "theta_min=", 10))        theta_min = atof(a + 10);
        else if (!strncmp(a, "theta_max=", 10))        theta_max = atof(a + 10);
        else if (!strncmp(a, "bnc_min=", 8))           bnc_min = atoi(a + 8);
        else if (!strncmp(a, "bnc_max=", 8))           bnc_max = atoi(a + 8);
        else if (!strncmp(a, "bounces=", 8))           bnc_min = bnc_max = atoi(a + 8);
#ifdef REF_GLOB
        else if (!strncmp(a, "lat_src=", 8))           Source_Loc[0] = atof(a + 8);
        else if (!strncmp(a, "lon_src=", 8))           Source_Loc[1] = atof(a + 8);
        else if (!strncmp(a, "lat_rcvr=", 9))          Receiver_Loc[0] = atof(a + 9);
        else if (!strncmp(a, "lon_rcvr=", 9))          Receiver_Loc[1] = atof(a + 9);
#else
        else if (!strncmp(a, "x_src=", 6))             Source_Loc[0] = atof(a + 6);
        else if (!strncmp(a, "y_src=", 6))             Source_Loc[1] = atof(a + 6);
        else if (!strncmp(a, "x_rcvr=", 7))            Receiver_Loc[0] = atof(a + 7);
        else if (!strncmp(a, "y_rcvr=", 7))            Receiver_Loc[1] = atof(a + 7);
#endif
        else if (!strncmp(a, "z_src=", 6))             Source_Loc[2] = atof(a + 6);
        else if (!strncmp(a, "direct=", 7))            direct = atoi(a + 7);
        else if (!strncmp(a, "theta_est=", 10))        theta_direct = atof(a + 10);
        else if (!strncmp(a, "phi_est=", 8))           { phi_direct = 90.0 - atof(a + 8); have_phi = true; }      // GeoAc3D_main.cpp:588
        else if (!strncmp(a, "verbose=", 8))           verbose_output = atoi(a + 8) != 0;
        else if (!strncmp(a, "azimuth_err_lim=", 16))  azimuth_err_lim = atof(a + 16);
        else if (!strncmp(a, "iterations=", 11))       iterations = atof(a + 11);
        else if (!strncmp(a, "freq=", 5))              freq = atof(a + 5);
        else if (!strncmp(a, "abs_coeff=", 10))        tweak_abs = std::max(0.0, atof(a + 10));
        else if (!strncmp(a, "z_grnd=", 7))            z_grnd = atof(a + 7);
        else if (!strncmp(a, "alt_max=", 8))           GeoAc_vert_limit = atof(a + 8);
#ifndef REF_RNGDEP
        else if (!strncmp(a, "rng_max=", 8))           GeoAc_range_limit = atof(a + 8);
#endif
        else if (!strncmp(a, "profile_format=", 15)) {}
        else { fprintf(stderr, "unknown key %s\n", a); return 2; }
    }
    Source_Loc[2] = std::max(z_grnd, Source_Loc[2]);
#ifdef REF_RNGDEP
    Spline_Multi_G2S((char*)prof[0].c_str(), (char*)prof[1].c_str(), (char*)prof[2].c_str(), fmt);
    GeoAc_SetPropRegion();
#elif defined(REF_GLOBAL)
    Spline_Single_G2S((char*)prof[0].c_str(), fmt);               // GeoAcGlobal_main.cpp:547 loads the profile a second time
#endif
    char title[8] = "e";
    results.open("e_results.dat");
    std::vector<double> rows;
    const auto t0 = std::chrono::steady_clock::now();
    if (direct) {                                      // GeoAc3D_RunEigDirect, GeoAc3D_main.cpp:546-601 (bounces= sets bnc_min = bnc_max here)
#ifndef REF_GLOB
        if (!have_phi) phi_direct = 180.0 / 3.14159 * atan2(Receiver_Loc[1], Receiver_Loc[0]);                       // :586
#endif
        double th = theta_direct, ph = phi_direct;
        const int before = eigenray_count;
        GeoAc_3DEigenray_LM(Source_Loc, Receiver_Loc, th, ph, freq, bnc_min, iterations, title);
        double row[8] = { (double)bnc_min, 1.0, theta_direct, phi_direct, 0.0, eigenray_count > before ? 1.0 : 0.0, th, ph };
        rows.insert(rows.end(), row, row + 8);
    } else
    for (int n_bnc = bnc_min; n_bnc <= bnc_max; n_bnc++) {
        double theta_start = theta_min, theta_next, theta_est, phi_est;
        while (theta_start < theta_max) {
            const bool ok = GeoAc_EstimateEigenray(Source_Loc, Receiver_Loc, theta_start, theta_max, theta_est, phi_est, theta_next, n_bnc, azimuth_err_lim);
            double row[8] = { (double)n_bnc, ok ? 1.0 : 0.0, theta_est, ok ? phi_est : 0.0, theta_next, 0.0, 0.0, 0.0 };
            if (ok) {
                const int before = eigenray_count;
                GeoAc_3DEigenray_LM(Source_Loc, Receiver_Loc, theta_est, phi_est, freq, n_bnc, iterations, title);
                row[5] = eigenray_count > before ? 1.0 : 0.0; row[6] = theta_est; row[7] = phi_est;
            }
            rows.insert(rows.end(), row, row + 8);
            theta_start = theta_next;
        }
    }
    results.close();
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    FILE* f = fopen(out_path.c_str(), "wb");
    if (!f) { perror("out"); return 2; }
    const double hdr[2] = { (double)(rows.size() / 8), secs };
    fwrite(hdr, sizeof(double), 2, f);
    fwrite(rows.data(), sizeof(double), rows.size(), f);
    fclose(f);
    fprintf(stderr, "ref_eig: %zu estimate calls, %d eigenrays, %.2f s\n", rows.size() / 8, eigenray_count, secs);
    return 0;
}
