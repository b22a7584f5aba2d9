/* Stand-in for <gsl/gsl_permutation.h> (see gsl_rng.h in this directory). */
#ifndef GTB200_GSL_STUB_PERMUTATION_H
#define GTB200_GSL_STUB_PERMUTATION_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct { size_t size; size_t *data; } gsl_permutation;
#ifdef __cplusplus
}
#endif
#endif
