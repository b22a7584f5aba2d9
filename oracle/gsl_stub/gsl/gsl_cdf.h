/* Stand-in for <gsl/gsl_cdf.h> (see gsl_rng.h in this directory). */
#ifndef GTB200_GSL_STUB_CDF_H
#define GTB200_GSL_STUB_CDF_H
#ifdef __cplusplus
extern "C" {
#endif
double gsl_cdf_binomial_Q(unsigned int k, double p, unsigned int n);
double gsl_cdf_binomial_P(unsigned int k, double p, unsigned int n);
double gsl_cdf_poisson_Q(unsigned int k, double mu);
double gsl_cdf_poisson_P(unsigned int k, double mu);
double gsl_cdf_tdist_Q(double x, double nu);
double gsl_cdf_tdist_P(double x, double nu);
double gsl_cdf_ugaussian_Q(double x);
double gsl_cdf_ugaussian_P(double x);
double gsl_cdf_hypergeometric_Q(unsigned int k, unsigned int n1, unsigned int n2, unsigned int t);
double gsl_cdf_hypergeometric_P(unsigned int k, unsigned int n1, unsigned int n2, unsigned int t);
#ifdef __cplusplus
}
#endif
#endif
