/* Test-only host build of iak3d_b200/csrc/iak3d_math.h: lets the CPU suite compare the arithmetic the
 * device table kernels run (same source) with the oracle.  Never loaded by the product. */
#include "../iak3d_b200/csrc/iak3d_math.h"

extern "C" {

int hm_avcov(double a, double b, double c, double d, double ad, double nud, int sdfdType, double s1, double s2, double* out) {
  IaPrm P;
  int st = iak_prm_init(&P, ad, nud, sdfdType, s1, s2);
  if (st != 0) return st;
  *out = iak_avcov(a, b, c, d, &P);
  return 0;
}
int hm_avvar(double a, double b, double ad, double nud, int sdfdType, double s1, double s2, double* out) {
  IaPrm P;
  int st = iak_prm_init(&P, ad, nud, sdfdType, s1, s2);
  if (st != 0) return st;
  *out = iak_avvar(a, b, &P);
  return 0;
}
double hm_besselk(double nu, double x) { return iak_besselk(nu, x, nullptr); }
/* the table-driven variant the device kernels use */
double hm_besselk_rt(double nu, double x) {
  double rt[4 * IAK_BK_NRT];
  for (int i = 1; i < IAK_BK_NRT; i++) iak_besselk_table(nu, rt, i);
  return iak_besselk(nu, x, rt);
}
/* the Chebyshev-fit variant the device table kernels use for x >= 2 (order-dependent work done once, hx_init) */
double hm_besselk_fit(double nu, double x) {
  double rt[4 * IAK_BK_NRT];
  for (int i = 1; i < IAK_BK_NRT; i++) iak_besselk_table(nu, rt, i);
  BkFit F;
  iak_bk_fit_init(&F, nu);
  return iak_besselk(nu, x, rt, &F);
}
/* Matern / spherical only (the Wendland coefficients are prepared in tables.cu) */
double hm_phix(double D, int modelx, double a, double nu) {
  HxPrm H;
  H.modelx = modelx; H.a = a; H.nu = nu;
  H.s = sqrt(2.0 * nu) / a;
  H.lconst = -(nu - 1.0) * log(2.0) - lgamma(nu);
  H.half = (nu == 0.5) ? 1 : (nu == 1.5) ? 3 : (nu == 2.5) ? 5 : 0;
  H.bk.on = 0;
  if (H.half == 0) iak_bk_fit_init(&H.bk, nu);
  return iak_phix(D, &H);
}
}
