/* oracle/orc_spline1d.c -- TEST INFRASTRUCTURE (CPU oracle).
 * 1-D natural cubic spline in knot-slope (Hermite) form: restates Code/Atmo/G2S_Spline1D.cpp:161-281
 * (identical code in G2S_GlobalSpline1D.cpp:175-296) keeping its floating-point expression order.
 */
#include <math.h>
#include <stdlib.h>
#include "geoac_oracle.h"

/* Thomas algorithm for the knot slopes of a natural spline -- G2S_Spline1D.cpp:161-196 */
void orc_spline1d_set_slopes(orc_spline1d* s) {
    const int n = s->n;
    const double* x = s->x; const double* f = s->f;
    double* cp = (double*)malloc(sizeof(double) * (size_t)n);
    double* dp = (double*)malloc(sizeof(double) * (size_t)n);
    double lower, diag, upper, rhs;

    diag  = 2.0 / (x[1] - x[0]);
    upper = 1.0 / (x[1] - x[0]);
    rhs   = 3.0 * (f[1] - f[0]) / pow(x[1] - x[0], 2);
    cp[0] = upper / diag;
    dp[0] = rhs / diag;

    for (int i = 1; i < n - 1; i++) {
        lower = 1.0 / (x[i] - x[i - 1]);
        diag  = 2.0 * (1.0 / (x[i] - x[i - 1]) + 1.0 / (x[i + 1] - x[i]));
        upper = 1.0 / (x[i + 1] - x[i]);
        rhs   = 3.0 * ((f[i] - f[i - 1]) / pow(x[i] - x[i - 1], 2)
                     + (f[i + 1] - f[i]) / pow(x[i + 1] - x[i], 2));
        cp[i] = upper / (diag - cp[i - 1] * lower);
        dp[i] = (rhs - dp[i - 1] * lower) / (diag - cp[i - 1] * lower);
    }

    lower = 1.0 / (x[n - 1] - x[n - 2]);
    diag  = 2.0 / (x[n - 1] - x[n - 2]);
    rhs   = 3.0 * (f[n - 1] - f[n - 2]) / pow(x[n - 1] - x[n - 2], 2);
    dp[n - 1] = (rhs - dp[n - 2] * lower) / (diag - cp[n - 2] * lower);

    s->slopes[n - 1] = dp[n - 1];
    for (int i = n - 2; i > -1; i--) s->slopes[i] = dp[i] - cp[i] * s->slopes[i + 1];
    free(cp); free(dp);
}

/* cursor-accelerated interval search -- G2S_Spline1D.cpp:202-243.  Callers clamp x into [x0, x_{n-1}] first. */
static int find_segment(double x, const double* xs, int n, int* cursor) {
    int prev = *cursor;
    int hit = 0;
    if (x >= xs[prev] && x <= xs[prev + 1]) hit = 1;
    if (!hit && prev + 2 <= n - 1) {
        if (x >= xs[prev + 1] && x <= xs[prev + 2]) { hit = 1; prev = prev + 1; }
    }
    if (!hit && prev - 1 >= 0) {
        if (x >= xs[prev - 1] && x <= xs[prev]) { hit = 1; prev = prev - 1; }
    }
    if (!hit) {
        int index = n + 1;
        for (int i = 0; i < n; i++) {           /* two-ended linear scan, same visiting order as the reference */
            if (i + 1 < n && x >= xs[i] && x <= xs[i + 1]) { index = i; break; }
            if (n - 2 - i >= 0 && x >= xs[n - 2 - i] && x < xs[n - 1 - i]) { index = (n - 2) - i; break; }
        }
        prev = index;
    }
    *cursor = prev;
    return prev;
}

double orc_spline1d_f(double x, orc_spline1d* s) {
    int k = find_segment(x, s->x, s->n, &s->accel);
    if (k >= s->n) return 0.0;
    double X = (x - s->x[k]) / (s->x[k + 1] - s->x[k]);
    double A = s->slopes[k] * (s->x[k + 1] - s->x[k]) - (s->f[k + 1] - s->f[k]);
    double B = -s->slopes[k + 1] * (s->x[k + 1] - s->x[k]) + (s->f[k + 1] - s->f[k]);
    return (1.0 - X) * s->f[k] + X * s->f[k + 1] + X * (1.0 - X) * (A * (1.0 - X) + B * X);
}

double orc_spline1d_df(double x, orc_spline1d* s) {
    int k = find_segment(x, s->x, s->n, &s->accel);
    if (k >= s->n) return 0.0;
    double X = (x - s->x[k]) / (s->x[k + 1] - s->x[k]);
    double A = s->slopes[k] * (s->x[k + 1] - s->x[k]) - (s->f[k + 1] - s->f[k]);
    double B = -s->slopes[k + 1] * (s->x[k + 1] - s->x[k]) + (s->f[k + 1] - s->f[k]);
    return (s->f[k + 1] - s->f[k]) / (s->x[k + 1] - s->x[k])
         + (1.0 - 2.0 * X) * (A * (1.0 - X) + B * X) / (s->x[k + 1] - s->x[k])
         + X * (1.0 - X) * (B - A) / (s->x[k + 1] - s->x[k]);
}

double orc_spline1d_ddf(double x, orc_spline1d* s) {
    int k = find_segment(x, s->x, s->n, &s->accel);
    if (k >= s->n) return 0.0;
    double X = (x - s->x[k]) / (s->x[k + 1] - s->x[k]);
    double A = s->slopes[k] * (s->x[k + 1] - s->x[k]) - (s->f[k + 1] - s->f[k]);
    double B = -s->slopes[k + 1] * (s->x[k + 1] - s->x[k]) + (s->f[k + 1] - s->f[k]);
    return 2.0 * (B - 2.0 * A + (A - B) * 3.0 * X) / pow(s->x[k + 1] - s->x[k], 2);
}
