/* test-only export of lsd_math.h for ctypes (tests/test_math.py) */
#include "lsd_math.h"
double t_sin(double x) { return lsdm_sin(x); }
double t_cos(double x) { return lsdm_cos(x); }
double t_atan2(double y, double x) { return lsdm_atan2(y, x); }
double t_atan(double x) { return lsdm_atan(x); }
double t_exp(double x) { return lsdm_exp(x); }
double t_log(double x) { return lsdm_log(x); }
double t_log10(double x) { return lsdm_log10(x); }
double t_sinh(double x) { return lsdm_sinh(x); }
double t_pow(double x, double y) { return lsdm_pow(x, y); }
/* phase-1 parts, to measure the fast-path error bound */
void t_sincos_fast(double x, int want_cos, double* h, double* l) {
    lsdm_dd r; int n = (lsdm_rem_pio2(x, &r) + want_cos) & 3;
    lsdm_dd f = lsdm_sincos_fast(r.h, r.l, n & 1);
    if (n & 2) { f.h = -f.h; f.l = -f.l; }
    *h = f.h; *l = f.l;
}
void t_atan2_fast(double y, double x, double* h, double* l) {
    double ay = fabs(y), ax = fabs(x); int swap = ay > ax, xneg = x < 0;
    lsdm_dd f = swap ? lsdm_atan_ratio_fast(ax, ay) : lsdm_atan_ratio_fast(ay, ax);
    if (swap) f = lsdm_const_minus(LSDM_PIO2_H, LSDM_PIO2_L, f);
    if (xneg) f = lsdm_const_minus(LSDM_PI_H, LSDM_PI_L, f);
    *h = f.h; *l = f.l;
}
/* phase-2 only */
double t_sin_slow(double x) { lsdm_dd r; int n = lsdm_rem_pio2(x, &r) & 3; return lsdm_sincos_slow(r, n); }
double t_cos_slow(double x) { lsdm_dd r; int n = (lsdm_rem_pio2(x, &r) + 1) & 3; return lsdm_sincos_slow(r, n); }
double t_atan2_slow(double y, double x) {
    double ay = fabs(y), ax = fabs(x); double r = lsdm_atan2_slow(ay, ax, ay > ax, x < 0); return y < 0 ? -r : r;
}
/* vectorised drivers */
void v_sin(const double* x, double* o, int n) { for (int i = 0; i < n; i++) o[i] = lsdm_sin(x[i]); }
void v_cos(const double* x, double* o, int n) { for (int i = 0; i < n; i++) o[i] = lsdm_cos(x[i]); }
void v_atan2(const double* y, const double* x, double* o, int n) { for (int i = 0; i < n; i++) o[i] = lsdm_atan2(y[i], x[i]); }
void v_exp(const double* x, double* o, int n) { for (int i = 0; i < n; i++) o[i] = lsdm_exp(x[i]); }
void v_log(const double* x, double* o, int n) { for (int i = 0; i < n; i++) o[i] = lsdm_log(x[i]); }
void v_log10(const double* x, double* o, int n) { for (int i = 0; i < n; i++) o[i] = lsdm_log10(x[i]); }
void v_sinh(const double* x, double* o, int n) { for (int i = 0; i < n; i++) o[i] = lsdm_sinh(x[i]); }
void v_pow(const double* x, const double* y, double* o, int n) { for (int i = 0; i < n; i++) o[i] = lsdm_pow(x[i], y[i]); }
/* the phase-1-only entry points the stencil stage uses: out = value where ok, count of ok returned */
int v_atan2_try(const double* y, const double* x, double* o, unsigned char* ok, int n) {
    int k = 0;
    for (int i = 0; i < n; i++) { ok[i] = (unsigned char)lsdm_atan2_try(y[i], x[i], &o[i]); k += ok[i]; }
    return k;
}
int v_sincos_try(const double* x, double* s, double* c, unsigned char* ok, int n) {
    int k = 0;
    for (int i = 0; i < n; i++) { ok[i] = (unsigned char)lsdm_sincos_try(x[i], &s[i], &c[i]); k += ok[i]; }
    return k;
}
