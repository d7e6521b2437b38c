/* TEST INFRASTRUCTURE.  Oracle (ii): binds every libm call of the unmodified reference
 * (built -fno-builtin, linked -Bsymbolic) to the repo's portable math, so that the reference's
 * own control flow runs on exactly the arithmetic the device uses.  sqrt/floor/ceil/round stay
 * libc's (IEEE-exact on both sides).  SURVEY.md §8c "two-oracle strategy". */
#include "lsd_math.h"
double sin(double x) { return lsdm_sin(x); }
double cos(double x) { return lsdm_cos(x); }
double atan2(double y, double x) { return lsdm_atan2(y, x); }
double atan(double x) { return lsdm_atan(x); }
double exp(double x) { return lsdm_exp(x); }
double log(double x) { return lsdm_log(x); }
double log10(double x) { return lsdm_log10(x); }
double sinh(double x) { return lsdm_sinh(x); }
double pow(double x, double y) { return lsdm_pow(x, y); }
