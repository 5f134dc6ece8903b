// gemm.cu -- dense FP64 products behind the explicit inverse (chol2inv / solve(C): optim_functions.R:297,
// fitIAK3D.R:931) and the Newton-Raphson terms (gradHessIAK3D, nrUpdatesIAK3D.R:118-216).
#include "common.cuh"

int iak_inverse_from_factor(iak3d_ctx* ctx, Slot* s, double* d_iC) {
  (void)s; (void)d_iC;
  ctx->err = "explicit inverse: not implemented yet";
  return IAK3D_ERR_ARG;
}

extern "C" int iak3d_gradhess(iak3d_ds* ds, const iak3d_pars* pars, int32_t ncomp, const double* lnvPars, double* nll, double* grad,
                              double* hess, double* fim, double* betahat, double* vbetahat) {
  (void)pars; (void)ncomp; (void)lnvPars; (void)nll; (void)grad; (void)hess; (void)fim; (void)betahat; (void)vbetahat;
  if (!ds) return IAK3D_ERR_ARG;
  ds->ctx->err = "gradhess: not implemented yet";
  return IAK3D_ERR_ARG;
}
