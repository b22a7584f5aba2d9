/* Link-time stand-ins for the handful of GSL symbols the reference objects reference.
 * TEST INFRASTRUCTURE ONLY (oracle build). The overlap/count/scan path never calls GSL
 * arithmetic; genomic_scans only allocates+seeds a generator at start-up. Anything
 * statistical aborts loudly so a wrong result can never be produced silently. */
#include <stdio.h>
#include <stdlib.h>
#include "gsl/gsl_rng.h"
#include "gsl/gsl_cdf.h"
#include "gsl/gsl_randist.h"

static const gsl_rng_type stub_type = { "gtb200-stub" };
const gsl_rng_type *gsl_rng_default = &stub_type;

static void die(const char *what) {
  fprintf(stderr, "[gsl stub] %s is not available in the oracle build\n", what);
  abort();
}
const gsl_rng_type *gsl_rng_env_setup(void) { return gsl_rng_default; }
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T) {
  gsl_rng *r = (gsl_rng *)malloc(sizeof(gsl_rng));
  r->type = T; r->state = 0; return r;
}
void gsl_rng_set(const gsl_rng *r, unsigned long seed) { ((gsl_rng *)r)->state = seed; }
void gsl_rng_free(gsl_rng *r) { free(r); }
unsigned long gsl_rng_uniform_int(const gsl_rng *r, unsigned long n) { (void)r; (void)n; die("gsl_rng_uniform_int"); return 0; }
double gsl_cdf_binomial_Q(unsigned int k, double p, unsigned int n) { (void)k; (void)p; (void)n; die("gsl_cdf_binomial_Q"); return 0; }
double gsl_cdf_binomial_P(unsigned int k, double p, unsigned int n) { (void)k; (void)p; (void)n; die("gsl_cdf_binomial_P"); return 0; }
double gsl_cdf_poisson_Q(unsigned int k, double mu) { (void)k; (void)mu; die("gsl_cdf_poisson_Q"); return 0; }
double gsl_cdf_poisson_P(unsigned int k, double mu) { (void)k; (void)mu; die("gsl_cdf_poisson_P"); return 0; }
double gsl_cdf_tdist_Q(double x, double nu) { (void)x; (void)nu; die("gsl_cdf_tdist_Q"); return 0; }
double gsl_cdf_tdist_P(double x, double nu) { (void)x; (void)nu; die("gsl_cdf_tdist_P"); return 0; }
double gsl_cdf_ugaussian_Q(double x) { (void)x; die("gsl_cdf_ugaussian_Q"); return 0; }
double gsl_cdf_ugaussian_P(double x) { (void)x; die("gsl_cdf_ugaussian_P"); return 0; }
double gsl_cdf_hypergeometric_Q(unsigned int k, unsigned int n1, unsigned int n2, unsigned int t) { (void)k; (void)n1; (void)n2; (void)t; die("gsl_cdf_hypergeometric_Q"); return 0; }
double gsl_cdf_hypergeometric_P(unsigned int k, unsigned int n1, unsigned int n2, unsigned int t) { (void)k; (void)n1; (void)n2; (void)t; die("gsl_cdf_hypergeometric_P"); return 0; }
unsigned int gsl_ran_poisson(const gsl_rng *r, double mu) { (void)r; (void)mu; die("gsl_ran_poisson"); return 0; }
unsigned int gsl_ran_binomial(const gsl_rng *r, double p, unsigned int n) { (void)r; (void)p; (void)n; die("gsl_ran_binomial"); return 0; }
void gsl_ran_shuffle(const gsl_rng *r, void *base, size_t nmembm, size_t size) { (void)r; (void)base; (void)nmembm; (void)size; die("gsl_ran_shuffle"); }
double gsl_ran_binomial_pdf(unsigned int k, double p, unsigned int n) { (void)k; (void)p; (void)n; die("gsl_ran_binomial_pdf"); return 0; }
double gsl_ran_poisson_pdf(unsigned int k, double mu) { (void)k; (void)mu; die("gsl_ran_poisson_pdf"); return 0; }
