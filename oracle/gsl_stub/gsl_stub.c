/* Link-time stand-ins for the handful of GSL symbols the reference objects reference.
 * TEST INFRASTRUCTURE ONLY (oracle build). The overlap/count/scan path never calls GSL
 * arithmetic; genomic_scans only allocates+seeds a generator at start-up.
 * The three tail probabilities `genomic_scans peaks` needs (binomial, Poisson, normal) are
 * given by their defining sums in long double -- a formulation independent of the continued
 * fractions the product uses (csrc/gtb_scan.cu) -- so that the reference's PeakFinder can run
 * here as the oracle of the peaks driver (GSL itself is not in this image: parity with GSL's
 * own digits is unpinned).  Anything else statistical aborts loudly so a wrong
 * result can never be produced silently. */
#include <math.h>
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
/* P(X > k), X ~ Bin(n, p): the sum of the probability masses k+1 .. n, largest first */
double gsl_cdf_binomial_Q(unsigned int k, double p, unsigned int n) {
  if (p > 1.0 || p < 0.0) die("gsl_cdf_binomial_Q (p outside [0,1])");
  if (k >= n) return 0.0;
  if (p == 0.0) return 0.0;
  if (p == 1.0) return 1.0;
  const long double lp = logl((long double)p), lq = log1pl(-(long double)p), lgn = lgammal((long double)n + 1.0L);
  unsigned int mode = (unsigned int)floorl(((long double)n + 1.0L) * p);
  if (mode > n) mode = n;
  long double sum = 0.0L;
  /* upwards from max(k+1, mode), then downwards from there to k+1 */
  unsigned int from = k + 1 > mode ? k + 1 : mode;
  for (unsigned int i = from; i <= n; i++) {
    const long double t = expl(lgn - lgammal((long double)i + 1.0L) - lgammal((long double)(n - i) + 1.0L) + i * lp + (n - i) * lq);
    sum += t;
    if (t < sum * 1e-30L && i > mode) break;
  }
  for (unsigned int i = from; i-- > k + 1;) {
    const long double t = expl(lgn - lgammal((long double)i + 1.0L) - lgammal((long double)(n - i) + 1.0L) + i * lp + (n - i) * lq);
    sum += t;
    if (t < sum * 1e-30L) break;
  }
  return (double)(sum > 1.0L ? 1.0L : sum);
}
double gsl_cdf_binomial_P(unsigned int k, double p, unsigned int n) { (void)k; (void)p; (void)n; die("gsl_cdf_binomial_P"); return 0; }
/* P(X > k), X ~ Poisson(mu) */
double gsl_cdf_poisson_Q(unsigned int k, double mu) {
  if (mu <= 0.0) die("gsl_cdf_poisson_Q (mu <= 0)");
  const long double lmu = logl((long double)mu);
  unsigned int mode = (unsigned int)floor(mu);
  unsigned int from = k + 1 > mode ? k + 1 : mode;
  long double sum = 0.0L;
  for (unsigned int i = from;; i++) {
    const long double t = expl(-(long double)mu + i * lmu - lgammal((long double)i + 1.0L));
    sum += t;
    if ((t < sum * 1e-30L && i > mode) || i > from + 100000u) break;
  }
  for (unsigned int i = from; i-- > k + 1;) {
    const long double t = expl(-(long double)mu + i * lmu - lgammal((long double)i + 1.0L));
    sum += t;
    if (t < sum * 1e-30L) break;
  }
  return (double)(sum > 1.0L ? 1.0L : sum);
}
double gsl_cdf_poisson_P(unsigned int k, double mu) { (void)k; (void)mu; die("gsl_cdf_poisson_P"); return 0; }
double gsl_cdf_tdist_Q(double x, double nu) { (void)x; (void)nu; die("gsl_cdf_tdist_Q"); return 0; }
double gsl_cdf_tdist_P(double x, double nu) { (void)x; (void)nu; die("gsl_cdf_tdist_P"); return 0; }
double gsl_cdf_ugaussian_Q(double x) { return (double)(0.5L * erfcl((long double)x / sqrtl(2.0L))); }
double gsl_cdf_ugaussian_P(double x) { (void)x; die("gsl_cdf_ugaussian_P"); return 0; }
double gsl_cdf_hypergeometric_Q(unsigned int k, unsigned int n1, unsigned int n2, unsigned int t) { (void)k; (void)n1; (void)n2; (void)t; die("gsl_cdf_hypergeometric_Q"); return 0; }
double gsl_cdf_hypergeometric_P(unsigned int k, unsigned int n1, unsigned int n2, unsigned int t) { (void)k; (void)n1; (void)n2; (void)t; die("gsl_cdf_hypergeometric_P"); return 0; }
unsigned int gsl_ran_poisson(const gsl_rng *r, double mu) { (void)r; (void)mu; die("gsl_ran_poisson"); return 0; }
unsigned int gsl_ran_binomial(const gsl_rng *r, double p, unsigned int n) { (void)r; (void)p; (void)n; die("gsl_ran_binomial"); return 0; }
void gsl_ran_shuffle(const gsl_rng *r, void *base, size_t nmembm, size_t size) { (void)r; (void)base; (void)nmembm; (void)size; die("gsl_ran_shuffle"); }
double gsl_ran_binomial_pdf(unsigned int k, double p, unsigned int n) { (void)k; (void)p; (void)n; die("gsl_ran_binomial_pdf"); return 0; }
double gsl_ran_poisson_pdf(unsigned int k, double mu) { (void)k; (void)mu; die("gsl_ran_poisson_pdf"); return 0; }
