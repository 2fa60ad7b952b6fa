/* TEST INFRASTRUCTURE ONLY (oracle shim). */
#ifndef ORACLE_GSL_STATS_H
#define ORACLE_GSL_STATS_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
double gsl_stats_mean(const double data[], size_t stride, size_t n);
#ifdef __cplusplus
}
#endif
#endif
