/* Minimal stand-in for <gsl/gsl_rng.h>: GSL is not installed in this image.
 * TEST INFRASTRUCTURE ONLY - lets the unmodified reference sources compile so the
 * reference binaries can serve as the parity oracle (oracle/Makefile). No hot-path
 * function of the reference performs GSL arithmetic; the random/statistics entry
 * points abort if they are ever reached. */
#ifndef GTB200_GSL_STUB_RNG_H
#define GTB200_GSL_STUB_RNG_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct { const char *name; } gsl_rng_type;
typedef struct { const gsl_rng_type *type; unsigned long state; } gsl_rng;
extern const gsl_rng_type *gsl_rng_default;
const gsl_rng_type *gsl_rng_env_setup(void);
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_set(const gsl_rng *r, unsigned long seed);
void gsl_rng_free(gsl_rng *r);
unsigned long gsl_rng_uniform_int(const gsl_rng *r, unsigned long n);
#ifdef __cplusplus
}
#endif
#endif
