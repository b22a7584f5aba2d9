/* Stand-in for <gsl/gsl_randist.h> (see gsl_rng.h in this directory). */
#ifndef GTB200_GSL_STUB_RANDIST_H
#define GTB200_GSL_STUB_RANDIST_H
#include <stddef.h>
#include "gsl_rng.h"
#ifdef __cplusplus
extern "C" {
#endif
unsigned int gsl_ran_poisson(const gsl_rng *r, double mu);
unsigned int gsl_ran_binomial(const gsl_rng *r, double p, unsigned int n);
void gsl_ran_shuffle(const gsl_rng *r, void *base, size_t nmembm, size_t size);
double gsl_ran_binomial_pdf(unsigned int k, double p, unsigned int n);
double gsl_ran_poisson_pdf(unsigned int k, double mu);
#ifdef __cplusplus
}
#endif
#endif
