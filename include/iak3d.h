/* iak3d.h -- C ABI of libiak3d.so: the B200 (sm_100a) implementation of iak3d's data-parallel
 * hot path (covariance build -> REML / composite-likelihood evaluation -> block kriging).
 *
 * The reference (ortont/iak3d, 100 % R) has no FFI boundary of its own; this ABI is what an R
 * `.Call` / `.C` binding (or Python ctypes) binds for the functions listed below.  Each entry point
 * cites the reference interface (file:line in the reference) that it replaces.
 *
 * Conventions
 *  - every pointer is a HOST pointer owned by the caller (R vectors); the library copies in/out.
 *  - matrices are column-major (R native); doubles are IEEE FP64; ids are int32.
 *  - return value: 0 = OK; 1 = matrix not positive definite / non-finite (R glue maps this to
 *    badnll = 9E99, fitIAK3D.R:685,785-806); 2 = invalid parameters (reference returns C = NA,
 *    fitIAK3D.R:739,885); <0 = CUDA failure (R glue calls stop()).  Never abort()/exit().
 *  - the wire format of the covariance parameters is the NATURAL scale list `parsBTfmd`
 *    (output of readParsIAK3D, fitIAK3D.R:1405-1408); transforms stay on the host.
 *  - no CPU fallback exists: every compute entry point fails with a negative status when no
 *    sm_100-class device is usable.
 */
#ifndef IAK3D_H
#define IAK3D_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define IAK3D_API __attribute__((visibility("default")))
#else
#define IAK3D_API
#endif

#define IAK3D_OK 0
#define IAK3D_NOT_SPD 1
#define IAK3D_BAD_PARS 2
#define IAK3D_ERR_CUDA (-1)
#define IAK3D_ERR_ARG (-2)
#define IAK3D_ERR_TIMEOUT (-3)

/* modelx, fitIAK3D.R:978-984 */
#define IAK3D_MODELX_NUGGET 0
#define IAK3D_MODELX_SPHERICAL 1
#define IAK3D_MODELX_MATERN 2
#define IAK3D_MODELX_WENDLAND 3

/* largest number of internal knots of a spline sd function (sdfdType > 0, fitIAK3D.R:1490-1495) */
#define IAK3D_MAX_SPL 12
/* largest number of site-specific random-effect terms (colnames4ssre, iaCovMatern.R:440-460) */
#define IAK3D_MAX_SSRE 8

/* parsBTfmd + the model switches that select terms (fitIAK3D.R:1405-1408, 955-956). */
typedef struct iak3d_pars {
  int32_t modelx;          /* IAK3D_MODELX_* */
  int32_t cmeOpt;          /* informational; cme is used as given (0 when cmeOpt == 0) */
  int32_t sdfdType[3];     /* cd1, cxd0, cxd1: 0 stationary, -1 exponential sd(d), -9 component absent, k > 0 natural
                              spline with k internal knots: then EVERY component goes through the interval-averaged
                              sd approximation of setCApproxIAK3D[2] (fitIAK3D.R:959-962,1919-2233) */
  int32_t quirk_cd1_unscaled; /* 1 = reproduce fitIAK3D.R:1052 (non-stationary cd1 term added without cd1) */
  double ax, nux, ad, nud; /* nud must be 0.5, 1.5 or 2.5 (iaCovMatern.R:12-29) */
  double cx0, cx1, cd1, cxd0, cxd1, cme;
  double sdfdPars[3][2];   /* (tau1, tau2) of each non-stationary component (iaCovMatern.R:40-43) */
  double quirk_scale;      /* multiplier of the unscaled cd1 term when quirk_cd1_unscaled applies; 0 is read as 1.
                              1 while optimising and for Ckk (predictIAK3D.R:183 rebuilds from parsBTfmd); cxdhat for
                              the kept C = cxdhat * A of a fit (fitIAK3D.R:871) */
  double sdfdSpl[3][IAK3D_MAX_SPL]; /* spline coefficients of a component with sdfdType k > 0 (the k free parameters;
                              iasdfd multiplies the basis by c(1, sdfdPars), fitIAK3D.R:1458) */
  int32_t nssre;           /* site-specific random effects (parsBTfmd$cssre, fitIAK3D.R:1413): number of terms, 0 = none */
  int32_t reserved_;
  double cssre[IAK3D_MAX_SSRE]; /* their variances: C += cssre[k] * listStructs4ssre[[k]] (fitIAK3D.R:1091-1101, 1233-1240) */
} iak3d_pars;

typedef struct iak3d_ctx iak3d_ctx;   /* one per GPU: stream, workspaces, error text */
typedef struct iak3d_ds iak3d_ds;     /* a dataset resident in HBM == `setupMats` + W of the reference */
typedef struct iak3d_fit iak3d_fit;   /* a kept factorisation == the big matrices of `lmmFit` */

/* ---- context ------------------------------------------------------------------------------ */
IAK3D_API int iak3d_ctx_create(int device, iak3d_ctx** out);
IAK3D_API int iak3d_ctx_destroy(iak3d_ctx* ctx);
IAK3D_API const char* iak3d_last_error(iak3d_ctx* ctx);
IAK3D_API int iak3d_device_info(iak3d_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, int64_t* total_mem);
/* number of kernels this library launched on ctx since creation (bench.py's gpu_launches claim) */
IAK3D_API int64_t iak3d_launch_count(iak3d_ctx* ctx);

/* ---- dataset == setupIAK3D (iaCovMatern.R:413-548) ----------------------------------------
 * xy: n x 2, dI: n x 2 (already rounded to the cm by the caller, fitIAK3D.R:127), W = cbind(X, z)
 * (fitIAK3D.R:757): n x q with q = p + 1 (W may be NULL with q = 0 when only C is wanted).
 * Unique locations / intervals keep first-occurrence order (`!duplicated`, iaCovMatern.R:418-419). */
IAK3D_API int iak3d_dataset_create(iak3d_ctx* ctx, int64_t n, const double* xy, const double* dI,
                         const double* W, int32_t q, iak3d_ds** out);
IAK3D_API int iak3d_dataset_destroy(iak3d_ds* ds);
IAK3D_API int iak3d_dataset_set_W(iak3d_ds* ds, const double* W, int32_t q);
/* *same = 1 when W equals, bit for bit, what the last iak3d_dataset_set_W stored (optim() passes the same data to every
 * callback; the R glue asks before uploading again) */
IAK3D_API int iak3d_dataset_same_W(iak3d_ds* ds, const double* W, int32_t q, int32_t* same);
IAK3D_API int iak3d_dataset_info(iak3d_ds* ds, int64_t* n, int64_t* nxU, int64_t* ndIU, int32_t* q);
/* 0-based ids of each row's unique location / unique depth interval (Kx, Kd as id vectors) */
IAK3D_API int iak3d_dataset_ids(iak3d_ds* ds, int32_t* xid, int32_t* did);
/* setupMats$XsdfdSplineU_{cd1,cxd0,cxd1} (iaCovMatern.R:520-539): interval-averaged spline basis of component comp
 * (0 cd1, 1 cxd0, 2 cxd1) on the dataset's UNIQUE depth intervals (first-occurrence order), ndIU x ncol column-major,
 * ncol = sdfdType + 1.  Needed before any evaluation whose sdfdType[comp] > 0. */
IAK3D_API int iak3d_dataset_set_sdfd_spline(iak3d_ds* ds, int32_t comp, int32_t ncol, const double* XsplU);
/* Lognormal data (lnTfmdData = TRUE; nrUpdatesIAK3DlnN.R:11-16,194): column `col` (0-based) of W is overwritten on the
 * device, at every evaluation, by sigma2Vec - diag(C) of that evaluation's parameters -- the bias term of the
 * lognormal mean, which changes with the covariance parameters.  col < 0 switches this off (default). */
IAK3D_API int iak3d_dataset_set_lnN_column(iak3d_ds* ds, int32_t col);

/* Site-specific random effects (the lmer-type option, siteIDData given: iaCovMatern.R:440-460).  listStructs4ssre[[k]] is
 * block diagonal over sites with blocks X[site, cn_k] X[site, cn_k]'; it is never formed: the dataset keeps the site id of
 * every row (int32, equal ids = same site; the caller maps its labels) and the nssre columns Xs = X[, colnames4ssre]
 * (n x nssre column-major), and every covariance built from the dataset adds
 *     [site(i) == site(j)] * sum_k pars->cssre[k] Xs(i, k) Xs(j, k).
 * nssre = 0 clears.  sigma2Vec is not defined with ssre (the reference sets it to NA, fitIAK3D.R:1092): NaN. */
IAK3D_API int iak3d_dataset_set_ssre(iak3d_ds* ds, int32_t nssre, const int32_t* siteid, const double* Xs);

/* ---- depth / horizontal kernels on unique pairs ---------------------------------------------
 * iaCovMatern2 (iaCovMatern.R:107-191): out is n1 x n2 column-major; avVar (n1) may be NULL
 * (avVarVec of set 1, iaCovMatern.R:53-54).  sdfdType in {0,-1}. */
IAK3D_API int iak3d_iacov(iak3d_ctx* ctx, int64_t n1, const double* dI1, int64_t n2, const double* dI2,
                double ad, double nud, int32_t sdfdType, const double* sdfdPars,
                double* out, double* avVar);
/* spatialCovIAK3D (iaCovMatern.R:327-408) applied element-wise to m distances, pars = (c1,a,nu) */
IAK3D_API int iak3d_spatial_cov(iak3d_ctx* ctx, int64_t m, const double* D, int32_t modelx,
                      double c1, double a, double nu, double* out);

/* ---- covariance assembly --------------------------------------------------------------------
 * setCIAK3D (fitIAK3D.R:955-1112): C (n x n, full symmetric, may be NULL) and sigma2Vec (n, may
 * be NULL). */
IAK3D_API int iak3d_cov_build(iak3d_ds* ds, const iak3d_pars* pars, double* C, double* sigma2Vec);
/* diag(setCIAK3D(pars)$C) only (n): what setmuIAK3D needs of C (fitIAK3D.R:921, 1261-1267) */
IAK3D_API int iak3d_cov_diag(iak3d_ds* ds, const iak3d_pars* pars, double* diagC);
/* setCIAK3D2 (fitIAK3D.R:1117-1252): cross-covariance between set 1 (n1 rows given here) and the
 * dataset (set 2); out is n1 x n, no cme (fitIAK3D.R:1243). */
IAK3D_API int iak3d_cov_cross(iak3d_ds* ds2, const iak3d_pars* pars, int64_t n1, const double* xy1,
                    const double* dI1, double* out);
/* setCApproxIAK3D2 (fitIAK3D.R:2095-2233): the same with spline sd functions.  Xspl1[c] is the spline basis of
 * component c on the ROWS of set 1 (n1 x (sdfdType[c]+1), column-major; setupMats$XsdfdSplineU_* expanded by Kd), or
 * NULL for a component whose sdfdType is not > 0. */
IAK3D_API int iak3d_cov_cross_spl(iak3d_ds* ds2, const iak3d_pars* pars, int64_t n1, const double* xy1,
                        const double* dI1, const double* const* Xspl1, double* out);

/* setCIAK3D2 with site-specific random effects (siteIDData / XData of set 1 and siteIDData2 / XData2 of the dataset,
 * iaCovMatern.R:645-663; fitIAK3D.R:1233-1240): siteid1 (n1), Xs1 (n1 x nssre) describe the rows given here. */
IAK3D_API int iak3d_cov_cross_ssre(iak3d_ds* ds2, const iak3d_pars* pars, int64_t n1, const double* xy1, const double* dI1,
                                   const int32_t* siteid1, const double* Xs1, double* out);

/* The matrix the likelihood actually factorises: setCIAK3D (fitIAK3D.R:955-1112) as the FACTOR BUILD forms it --
 * `cov_factor_kernel` into the blocked factor buffer, the kernel the covariance roofline is quoted on -- copied back as a
 * full symmetric n x n column-major matrix (the lower triangle is what the kernel wrote, the upper triangle its mirror).
 * batched != 0 sends two identical jobs through the batch entry point of the kernel (the job-array instantiation the
 * forward-difference gradient and the composite likelihood use) and returns the second; 0 uses the single-job one.
 * Wrows (q x n column-major, may be NULL) receives the bordered W' rows as the kernel laid them out.  Parity tests only. */
IAK3D_API int iak3d_factor_build_fetch(iak3d_ds* ds, const iak3d_pars* pars, int32_t batched, double* A, double* Wrows);

/* ---- factorisation --------------------------------------------------------------------------
 * lndetANDinvCb (optim_functions.R:268-312, methodSolveTEST == 1): chol(C), lndetC = 2 sum log diag,
 * invCb = C^-1 b (n x nrhs) or C^-1 itself when b == NULL (then invCb is n x n).  Returns
 * IAK3D_NOT_SPD where the reference returns NA. */
IAK3D_API int iak3d_lndet_invCb(iak3d_ctx* ctx, int64_t n, const double* C, int32_t nrhs, const double* b,
                      double* lndetC, double* invCb);

/* ---- likelihood terms == nllIAK3D(..., forCompLik = TRUE) (fitIAK3D.R:730-819) ---------------
 * Builds A = setCIAK3D(pars), factorises it and returns lndetA and WiAW = t(W) A^-1 W (q x q).
 * iAW (n x q, A^-1 W; fitIAK3D.R:782-808) is optional (NULL to skip the back-substitution). */
IAK3D_API int iak3d_nll_terms(iak3d_ds* ds, const iak3d_pars* pars, double* lndetA, double* WiAW, double* iAW);

/* ---- the closed-form tail == the rest of nllIAK3D (fitIAK3D.R:821-880; compositeLikIAK3D.R:191-265) -------------------------
 * Host arithmetic, O(p^3): betahat = solve(XiAX, XiAz), resiAres, the profiled variance cxdhat = resiAres / (n - p) [REML] or
 * / n, and nll = 0.5 ((n - p) log 2pi + lndetC + lndetXiCX + resiCres) [REML] or 0.5 (n log 2pi + lndetC + resiCres).
 * WiAW is q x q column-major (q = p + 1), only its UPPER triangle is read (forceSymmetric, fitIAK3D.R:811).  n_for_cxd /
 * n_for_lndet <= 0 select the defaults; the composite likelihood passes its own counts (compositeLikIAK3D.R:206-240).
 * *ok = 0 and *nll = 9e99 (the reference's badnll) where XiAX is not positive definite.  Outputs other than nll / ok may be NULL. */
IAK3D_API int iak3d_reml_tail(const double* WiAW, double lndetA, int64_t n, int32_t p, int32_t useReml, double n_for_cxd, double n_for_lndet,
                              double* nll, double* betahat, double* cxdhat, double* lndetXiAX, double* resiAres, int32_t* ok);
/* ---- the optimiser's callbacks, natively ---------------------------------------------------------------------------------
 * The model switches and parameter bounds that readParsIAK3D needs (fitIAK3D.R:1332-1417; parBnds: setInitsIAK3D.R). */
typedef struct iak3d_model {
  int32_t modelx;          /* IAK3D_MODELX_* : 0 nugget, 1 spherical, 2 matern, 3 wendland */
  int32_t cmeOpt, prodSum, lnTfmdData;
  int32_t sdfdType[3];     /* cd1, cxd0, cxd1 */
  int32_t reserved_;
  double nud;
  double axBnds[2], nuxBnds[2], adBnds[2], tau2Bnds[2], sxd1Bnds[2];
  double maxd;
} iak3d_model;
/* readParsIAK3D: unconstrained vector -> parsBTfmd, in the reference's consumption order: ax [, nux unless spherical], ad,
 * [cx0, cx1 if prodSum], [cd1 unless sdfdType_cd1 == -9], cxd1 (logit; cxd0 = 1 - cxd1) or (cxd0, cxd1) if lnTfmdData, the sd-function
 * parameters of the three components, [cme if cmeOpt == 1], then exp() of whatever is left = cssre.  IAK3D_ERR_ARG when the
 * vector is too short or an option is unknown. */
IAK3D_API int iak3d_read_pars(const iak3d_model* model, const double* pars, int32_t npar, iak3d_pars* out);
/* nllIAK3D(pars, ...) for the optimiser: transform + device terms + tail; 9e99 where the reference returns badnll. */
IAK3D_API int iak3d_nll_vector(iak3d_ds* ds, const iak3d_model* model, const double* pars, int32_t npar, int32_t useReml, double* nll);
/* gradnllIAK3D (fitIAK3D.R:1776-1911): npar + 1 evaluations in one batched launch, forward differences with the reference's
 * failure conventions (:1812-1813).  nll0 (may be NULL) <- the value at pars. */
IAK3D_API int iak3d_gradnll_fd(iak3d_ds* ds, const iak3d_model* model, const double* pars, int32_t npar, double delta, int32_t useReml,
                               double* grad, double* nll0);

/* nllIAK3D(pars, ..., rtnAll = FALSE) in one call: iak3d_nll_terms on the device + the tail above.  *status = IAK3D_OK /
 * IAK3D_NOT_SPD / IAK3D_BAD_PARS (then *nll = 9e99, fitIAK3D.R:885).  betahat (p), cxdhat, lndetA, WiAW (q*q) may be NULL. */
IAK3D_API int iak3d_nll_reml(iak3d_ds* ds, const iak3d_pars* pars, int32_t useReml, double* nll, double* betahat, double* cxdhat,
                             double* lndetA, double* WiAW, int32_t* status);
/* The same for `count` independent (dataset, parameter) pairs in one launch: the composite-
 * likelihood block loop (compositeLikIAK3D.R:49-93,105-155; mcmapply over blocks) and the forward-
 * difference gradient (fitIAK3D.R:1865-1886; npar+1 evaluations).  status[i] is per evaluation;
 * lndetA[i], WiAW + i*q*q are outputs.  All datasets must share q. */
IAK3D_API int iak3d_nll_terms_batch(iak3d_ctx* ctx, int32_t count, iak3d_ds* const* ds, const iak3d_pars* pars,
                          double* lndetA, double* WiAW, int32_t* status);
/* split form for device-resident timing: enqueue on the context stream, then fetch (sync + D2H). */
IAK3D_API int iak3d_nll_terms_batch_enqueue(iak3d_ctx* ctx, int32_t count, iak3d_ds* const* ds, const iak3d_pars* pars);
IAK3D_API int iak3d_nll_terms_batch_fetch(iak3d_ctx* ctx, double* lndetA, double* WiAW, int32_t* status);

/* The exchange step of the composite likelihood made cheap (compositeLikIAK3D.R:63-64, Reduce('+') over the workers; 81-83):
 * the same batch, but the per-unit terms never travel to the host -- they are summed on the device, unit i with weight[i]
 * into group[i] (0 <= group[i] < ngroups: one group per parameter set, so that the npar + 1 composite-likelihood
 * evaluations of a forward-difference gradient are ONE launch and ONE exchange):
 *     out[g] = { sum_i w_i forceSymmetric(WiAW_i) (q*q, column-major; upper triangle kept, fitIAK3D.R:811),
 *                sum_i w_i lndetA_i,  number of units of the group that failed (not SPD / invalid parameters) }
 * in the fixed order of i (deterministic).  d_out is a DEVICE pointer to ngroups * (q*q + 2) doubles owned by the caller --
 * the buffer a multi-GPU caller hands to ncclAllReduce next (the one entry point whose pointer is not a host pointer);
 * h_out (may be NULL) receives a host copy; d_out may be NULL when only the host copy is wanted (single GPU).  The
 * context's stream is synchronised before returning. */
IAK3D_API int iak3d_nll_terms_batch_wsum(iak3d_ctx* ctx, int32_t count, iak3d_ds* const* ds, const iak3d_pars* pars,
                               const double* weight, const int32_t* group, int32_t ngroups, double* d_out, double* h_out);
/* The same over several GPUs from ONE host process: R is single-threaded and a CUDA context does not survive fork() (the
 * reference's mcmapply over block pairs, compositeLikIAK3D.R:54), so the library runs one host thread per device.  Device d
 * evaluates its count[d] units -- ds / pars / weight / group are the per-device lists concatenated in device order, the datasets
 * of device d created on ctxs[d] -- and the per-device sums are added on the host in device order.  h_out: ngroups * (q*q + 2). */
IAK3D_API int iak3d_nll_terms_batch_wsum_multi(int32_t ndev, iak3d_ctx* const* ctxs, const int32_t* count, iak3d_ds* const* ds,
                                               const iak3d_pars* pars, const double* weight, const int32_t* group, int32_t ngroups,
                                               int32_t q, double* h_out);
/* ---- the exchange step over NVLink peer memory (one process per GPU) -----------------------------------------------------
 * The all-reduce of the composite-likelihood sums (compositeLikIAK3D.R:63-64) moves ~700 bytes: its cost is the latency of the
 * collective.  Instead of a library call, every rank owns a small mailbox in device memory that its peers map through CUDA IPC;
 * one single-CTA kernel per rank publishes its values, waits for the peers' flags and adds all mailboxes in rank order (the same
 * bits on every rank).  iak3d_peer_create returns this rank's 64-byte IPC handle; the caller exchanges the handles of all ranks
 * (any side channel: torch.distributed all_gather, MPI, a file) and passes them, in rank order, to iak3d_peer_connect.
 * iak3d_peer_allreduce sums n <= max_doubles doubles in place at the DEVICE pointer d_buf (e.g. the d_out of
 * iak3d_nll_terms_batch_wsum) on the context's stream and synchronises it.  world <= 16. */
typedef struct iak3d_peer iak3d_peer;
IAK3D_API int iak3d_peer_create(iak3d_ctx* ctx, int32_t rank, int32_t world, int32_t max_doubles, iak3d_peer** out, void* handle_out);
IAK3D_API int iak3d_peer_connect(iak3d_peer* peer, const void* handles /* world x 64 bytes, rank order */);
IAK3D_API int iak3d_peer_allreduce(iak3d_peer* peer, double* d_buf, int32_t n, double* h_out /* host copy of the sums, or NULL */,
                                   double* us /* device time of the exchange, or NULL */);
IAK3D_API int iak3d_peer_destroy(iak3d_peer* peer);

/* bytes of device memory one more simultaneous evaluation of this dataset needs (a slot: factor buffer + tables), and the
 * device memory currently free -- for callers that size their batches (composite-likelihood gradient) */
IAK3D_API int iak3d_dataset_slot_bytes(iak3d_ds* ds, int64_t* slot_bytes, int64_t* free_bytes);

/* ---- analytic gradient of the profiled REML / ML objective (replaces the npar + 1 evaluations of gradnllIAK3D,
 * fitIAK3D.R:1865-1886) --------------------------------------------------------------------------
 * pars[0] = the parameters; pars[k], k = 1..npar = the parameters with the k-th UNCONSTRAINED parameter moved by
 * delta[k-1] and passed through the host's transforms (readParsIAK3D) -- the candidates of the forward-difference
 * gradient.  One factorisation + explicit inverse + one pass over the lower triangle return lndetA and WiAW of
 * pars[0] (so the caller has the function value too) and
 *     gsum[k-1] = sum_ij w_ij (A(pars[k]) - A(pars[0]))_ij / delta[k-1] ,   d nll / d theta_k = gsum[k-1] / 2 ,
 * w = A^-1 - [REML] A^-1 X (X'A^-1 X)^-1 X'A^-1 - u u' / cxdhat,  u = A^-1 (z - X betahat).  gsum[k-1] is NaN for a
 * candidate with invalid parameters (the reference's NA difference, set to 0 by the caller, fitIAK3D.R:1813). */
IAK3D_API int iak3d_nll_grad(iak3d_ds* ds, int32_t npar, const iak3d_pars* pars, const double* delta, int32_t useReml,
                   double* lndetA, double* WiAW, double* gsum);

/* ---- kept fit == the big matrices of lmmFit (fitIAK3D.R:919-936; predictIAK3D.R:110-139) -----
 * pars are the RESCALED parameters (cxdhat folded in, fitIAK3D.R:857-863), so C = setCIAK3D(pars).
 * betahat (p), vbetahat (p x p).  Computes and keeps L = chol(C), iCX, iCz_muhat on the device. */
IAK3D_API int iak3d_fit_create(iak3d_ds* ds, const iak3d_pars* pars, const double* betahat, const double* vbetahat,
                     iak3d_fit** out);
/* the same with the residual z - muhat given by the caller (n): lognormal data, where muhat = setmuIAK3D(...)
 * (fitIAK3D.R:1261-1267, 919-923) is not X betahat */
IAK3D_API int iak3d_fit_create_res(iak3d_ds* ds, const iak3d_pars* pars, const double* betahat, const double* vbetahat,
                         const double* z_muhat, iak3d_fit** out);
IAK3D_API int iak3d_fit_destroy(iak3d_fit* fit);
/* copies out what lmmFit carries: iCX (n x p), iCz_muhat (n), and optionally C and iC (n x n, NULL to skip) */
IAK3D_API int iak3d_fit_get(iak3d_fit* fit, double* iCX, double* iCz_muhat, double* C, double* iC);

/* ---- block kriging == per-batch body of predictIAK3D (predictIAK3D.R:179-208,229-255) --------
 * m prediction supports (location, depth interval) with design rows XMap (m x p, column-major).
 * z (m), v (m) = universal-kriging prediction and variance incl. beta uncertainty. */
IAK3D_API int iak3d_predict(iak3d_fit* fit, int64_t m, const double* xyMap, const double* dIMap, const double* XMap,
                  double* z, double* v);
/* device-resident variant for timing: inputs staged by iak3d_predict_stage, results by _fetch */
IAK3D_API int iak3d_predict_stage(iak3d_fit* fit, int64_t m, const double* xyMap, const double* dIMap, const double* XMap);
/* with spline sd functions: Xspl1[c] = basis of component c on the m map supports (m x (sdfdType[c]+1)) or NULL */
IAK3D_API int iak3d_predict_stage_spl(iak3d_fit* fit, int64_t m, const double* xyMap, const double* dIMap, const double* XMap,
                            const double* const* Xspl1);
/* with site-specific random effects: site ids (m) and Xs = XMap[, colnames4ssre] (m x nssre) of the map supports
 * (predictIAK3D.R:166-195: siteIDMapThis, XMapThis handed to setupIAK3D / setupIAK3D2) */
IAK3D_API int iak3d_predict_stage_ssre(iak3d_fit* fit, int64_t m, const double* xyMap, const double* dIMap, const double* XMap,
                             const int32_t* siteidMap, const double* XsMap);
IAK3D_API int iak3d_predict_enqueue(iak3d_fit* fit);
IAK3D_API int iak3d_predict_fetch(iak3d_fit* fit, double* z, double* v);
/* the pieces z and v are made of (predictIAK3D.R:183-208), each m long, any may be NULL: Ckk = diag(setCIAK3D(map)),
 * sigma2Veck (its sigma2Vec), CkhiCz_muhat = Ckh iC (z - muhat), CkhiCChk = rowSums((Ckh iC) * Ckh).  Lognormal
 * predictions (predictIAK3D.R:229-255, setmuIAK3D) and profilePredictIAK3D are assembled from these by the caller.
 * Call after iak3d_predict_enqueue, before or after iak3d_predict_fetch. */
IAK3D_API int iak3d_predict_fetch_terms(iak3d_fit* fit, double* Ckk, double* sigma2Veck, double* CkhiCz_muhat, double* CkhiCChk);
/* The kriging panel as the predictor builds it: setCIAK3D2 (fitIAK3D.R:1117-1252) of chunk `chunk` of the staged map supports
 * against the data, made by the product's panel kernels (`phix_rect_kernel` + `cov_panel_kernel`, or the odd-row-count
 * fallback) in the blocked layout and copied back as rows x n column-major (rows = supports of that chunk; *rows_out).
 * Call after iak3d_predict_stage.  Parity tests only. */
IAK3D_API int iak3d_predict_panel_fetch(iak3d_fit* fit, int64_t chunk, double* Ckh, int64_t cap_rows, int64_t* rows_out);

/* options of a kept fit, set before iak3d_predict_stage:
 *  IAK3D_FIT_KEEP_CKHICX   keep Ckh iC X (m x p) of every batch for iak3d_predict_fetch_CkhiCX (composite-likelihood block
 *                          prediction sums it over block pairs, compositeLikIAK3D.R:590-599);
 *  IAK3D_FIT_CROSS_SQUARE  build Ckh like a block of setCIAK3D of the joined supports (compositeLikIAK3D.R:540-552), i.e.
 *                          WITH the cd1 quirk of fitIAK3D.R:1052, instead of setCIAK3D2. */
#define IAK3D_FIT_KEEP_CKHICX 1
#define IAK3D_FIT_CROSS_SQUARE 2
IAK3D_API int iak3d_fit_set_option(iak3d_fit* fit, int32_t option, int32_t value);
IAK3D_API int iak3d_predict_fetch_CkhiCX(iak3d_fit* fit, double* CkhiCX /* m x p column-major */);

/* ---- fixed-effect design matrix of interval supports == makeXvX (makeXvX.R:1-708), lnTfmdData = FALSE -------------------
 * predictIAK3D builds X for every (depth, batch) with makeXvX from the covariates of the map cells (predictIAK3D.R:155,
 * :405); at 10^6 cells x 6 intervals that host-side matrix (and its upload) is what the kriging waits for.  A design object
 * holds modelX + optionsModelX + nDiscPts + XLims flattened into two arrays (the host mirror iak3d_b200/design.py and the R
 * glue build them):
 *   ispec: type (0 'mlr' makeXvX.R:212-353, 1 'cubist' :360-434), p, ncov, nspl (dSpline.* columns), spl_kind (0 splines::bs,
 *          1 makeXnsClampedUG0[,-1]: allKnotsd2X, cubist2XIAK3D.R:187-208), nint (internal knots), nDiscPts, has_lims, infBnds;
 *          mlr:    per column  kind (0 no depth, 1 d, 2 d2, 3 d3, 4 dSpline), covariate column (-1 const; kind 4: spline
 *                  index), position k among the K depth columns in the order id, idInt, id2, id2Int, id3, id3Int (-1 for the
 *                  others): the reference recycles XLims over t(X[, depth columns]) from its first element, so entry
 *                  (row i, k) is clamped by limit column (i K + k) mod p (makeXvX.R:327-328);
 *          cubist: pc (rule columns), nRules, dcol (column of dIMidPts), number of distinct committees, nbreaks, nblocks;
 *                  per rule  has_const, nconds, ncoef, per condition (variable, dir 0 '<=' 1 '>' 2 'in', set length), the
 *                  coefficient variables; per rule column its committee block (-1 none); per block its rule range [r0, r1)
 *   dspec: boundary knots b1, b2, the internal knots, XLims[1,], XLims[2,]; mlr: per column the factor level its dummy
 *          indicates (NaN = numeric); cubist: per condition its threshold followed by its category set, then the sorted
 *          distinct depth breaks (getAlldBreaks, cubist2XIAK3D.R:889-900).
 * Capacities: p <= 64, <= 14 internal knots, <= 64 rules, <= 64 depth breaks. */
typedef struct iak3d_design iak3d_design;
IAK3D_API int iak3d_design_create(iak3d_ctx* ctx, const int32_t* ispec, int64_t nispec, const double* dspec, int64_t ndspec,
                                  iak3d_design** out);
IAK3D_API int iak3d_design_destroy(iak3d_design* design);
IAK3D_API int iak3d_design_info(iak3d_design* design, int32_t* p, int32_t* ncov);
/* X (m x p column-major) of m supports: covs m x ncov column-major (NULL when ncov == 0), dI m x 2.  splmin / splmax (m x nspl,
 * may be NULL): mlr -- min / max of the non-zero point values of each spline column over the row's nDiscPts depths; cubist --
 * splmin receives the trapezoid values of the spline columns (what XLims is taken from before they are replaced,
 * makeXvX.R:406-430) */
IAK3D_API int iak3d_design_eval(iak3d_design* design, int64_t m, const double* covs, const double* dI, double* X,
                                double* splmin, double* splmax);
/* lnTfmdData = TRUE (makeXvX.R:246-290, 437-600): X = column means of the point designs at the support's nDiscPts depths, vXU =
 * their sample covariance (R's var()) on the pU columns iU (0-based) that vary inside a support -- (m pU) x pU column-major,
 * block i at rows i pU .. i pU + pU - 1 (the layout setmuIAK3D / calcbetavXbeta read, fitIAK3D.R:1261-1328).  Limits act on
 * the point values; pmin / pmax (m x p, may be NULL) = min / max of each column's non-zero point values per support (NaN when
 * all are zero), from which a caller in set-limits mode forms XLims (:274-276). */
IAK3D_API int iak3d_design_eval_lnN(iak3d_design* design, int64_t m, const double* covs, const double* dI, int32_t pU,
                                    const int32_t* iU, double* X, double* vXU, double* pmin, double* pmax);
/* The whole prediction grid in one staging (the loops of predictIAK3D.R:144-155): ncell map cells (xyCell ncell x 2,
 * covCell ncell x ncov) x nint depth intervals (dIInt nint x 2); support (i, c) is row i * ncell + c of z, v (zMap / vMap
 * are ndIMap x nxMap) and its design row is made on the device.  nxPerBatch (predictIAK3D's argument, 2000 by default; 0 =
 * one batch) only matters when finite XLims apply to an mlr design: the reference's limit recycling (makeXvX.R:327-328)
 * depends on a row's position inside its batch.  Then iak3d_predict_enqueue / _fetch as usual. */
IAK3D_API int iak3d_predict_stage_grid(iak3d_fit* fit, iak3d_design* design, int64_t ncell, const double* xyCell,
                                       const double* covCell, int64_t nint, const double* dIInt, int64_t nxPerBatch);
/* the design rows of the staged job as the predictor uses them (m x p column-major); parity tests */
IAK3D_API int iak3d_predict_fetch_X(iak3d_fit* fit, double* X);

/* ---- composite-likelihood block construction == setVoronoiBlocks & co (compositeLikIAK3D.R:803-1133) ------------------
 * Locations resident on the device for repeated nearest-centroid passes: D = xyDist(x, vcentres) (iaCovMatern.R:888-920) and
 * which.min per row (first minimum; NaN centroids never win).  This is the Voronoi assignment of setVoronoiBlocks (:887-888),
 * getBlocksFromLocations for the map cells (:803-812) and calcnPerCluster (:1033-1046), the objective balanceNeighbours2's
 * random search evaluates thousands of times (:1048-1133).  block (m, 0-based, may be NULL), counts (nB, may be NULL). */
typedef struct iak3d_points iak3d_points;
IAK3D_API int iak3d_points_create(iak3d_ctx* ctx, int64_t m, const double* xy /* m x 2 */, iak3d_points** out);
IAK3D_API int iak3d_points_destroy(iak3d_points* pts);
IAK3D_API int iak3d_points_nearest(iak3d_points* pts, int32_t nB, const double* vcentres /* nB x 2 */, int32_t* block, int64_t* counts);
/* greedy seeding of the block centroids (compositeLikIAK3D.R:840-884) from the nU unique locations (first-occurrence order);
 * host algorithm, no device needed.  nBlocks = floor(nU / nPerBlock) (<= cap_blocks); vcentres nBlocks x 2 column-major. */
IAK3D_API int iak3d_voronoi_seed(int64_t nU, const double* xU, int32_t nPerBlock, int32_t cap_blocks, int32_t* nBlocks,
                                 double* vcentres);
/* the pairs (ind1, ind2) of deldir(x, y)$dirsgs (compositeLikIAK3D.R:907-916; package deldir, not vendored by the reference):
 * blocks whose Dirichlet tiles share an edge inside deldir's default window (bounding box + 10 % per side).  pairs: cap x 2
 * ROW-major int32, i < j, sorted; host algorithm (Bowyer-Watson). */
IAK3D_API int iak3d_dirichlet_adjacency(int32_t nB, const double* vcentres, int32_t* pairs, int64_t cap_pairs, int64_t* npairs);

/* ---- Newton-Raphson terms == gradHessIAK3D (nrUpdatesIAK3D.R:118-216) -----------------------
 * Components dA_i are the unit-variance terms of setCIAK3D in the order cx0, cx1, cd1, cxd0, cxd1,
 * cme (ncomp = 5 or 6); lnvPars (ncomp) their log-variances.  Outputs: nll, grad (ncomp), hess and
 * fim (ncomp x ncomp), betahat (p), vbetahat (p x p). */
IAK3D_API int iak3d_gradhess(iak3d_ds* ds, const iak3d_pars* pars, int32_t ncomp, const double* lnvPars,
                   double* nll, double* grad, double* hess, double* fim, double* betahat, double* vbetahat);

/* The same over several GPUs (only worth it at n = 40 000): every rank holds the dataset and calls this with its
 * rank; the factorisation, the inverse and everything O(n^2) are done by every rank (T = iC - iCX (X'iCX)^-1 X'iC must be
 * whole wherever a tile of T dC_i is formed), while the ncomp products T dC_i and the ncomp (ncomp + 1) / 2 pair traces -- 12 n^3
 * of the 13 n^3 flops -- are cut by ownership of 128 x 128 tile pairs {(R, C), (C, R)}.  Outputs identical on every rank:
 * nll, betahat, vbetahat, zTdCTz (ncomp: z'T dC_i T z), qij (ncomp x ncomp: z'T dC_i T dC_j T z).  tr_part (ncomp + ncomp^2)
 * = this rank's part of tr(T dC_i) and of tr(T dC_i T dC_j): ncclAllReduce(sum) it, then (nrUpdatesIAK3D.R:197-210)
 *   grad_i = (tr_i - zTdCTz_i) / 2,  fim_ij = tr_ij / 2,  hess_ij = (-tr_ij + 2 qij_ij) / 2  (+ (tr_i - zTdCTz_i) / 2 on the diagonal). */
IAK3D_API int iak3d_gradhess_shard(iak3d_ds* ds, const iak3d_pars* pars, int32_t ncomp, const double* lnvPars, int32_t rank,
                                   int32_t world, double* nll, double* betahat, double* vbetahat, double* zTdCTz, double* qij,
                                   double* tr_part);

/* ---- device-resident dense matrices ------------------------------------------------------------
 * The Eidsvik composite-likelihood block predictor predMatsIAK3D_CLEV_ByBlock (compositeLikIAK3D.R:609-798) is dense
 * algebra on matrices as large as (map supports + block data): `%*%` (:717-779), chol2inv(chol(.)) (:725, :755, :763,
 * :788), sub-block indexing C0klAll[c(i1,i2), i0] (:680-731), 0.5 * (A + t(A)) (:722, :759, :784-785) and
 * colSums(A * B) (:795).  These entry points are those operations on matrices that stay on the GPU; the joined-support
 * covariance C0klAll (:671) is built there by iak3d_mat_cov_build.  Shapes are logical (unpadded); sub-blocks are
 * (row0, col0, nrows, ncols), zero-based.  All calls are synchronous on the context's stream. */
typedef struct iak3d_mat iak3d_mat;
IAK3D_API int iak3d_mat_create(iak3d_ctx* ctx, int64_t rows, int64_t cols, iak3d_mat** out);      /* zero-filled */
IAK3D_API int iak3d_mat_destroy(iak3d_mat* M);
/* matrices are allocated stream-ordered from a cached pool; this returns the cache to the driver */
IAK3D_API int iak3d_mat_trim(iak3d_ctx* ctx);
IAK3D_API int iak3d_mat_shape(const iak3d_mat* M, int64_t* rows, int64_t* cols);
/* host column-major (leading dimension ld) -> sub-block, and back */
IAK3D_API int iak3d_mat_set(iak3d_mat* M, int64_t r0, int64_t c0, int64_t nr, int64_t nc, const double* src, int64_t ld);
IAK3D_API int iak3d_mat_get(const iak3d_mat* M, int64_t r0, int64_t c0, int64_t nr, int64_t nc, double* dst, int64_t ld);
/* D[dr0+i, dc0+j] = beta * D[dr0+i, dc0+j] + alpha * S[sr0+i, sc0+j]  (transpose != 0: S[sr0+j, sc0+i]); D != S */
IAK3D_API int iak3d_mat_copy(iak3d_mat* D, int64_t dr0, int64_t dc0, const iak3d_mat* S, int64_t sr0, int64_t sc0,
                             int64_t nr, int64_t nc, int32_t transpose, double alpha, double beta);
/* C = (accumulate ? C : 0) + alpha * A %*% t(B);  A: rows(C) x K, B: cols(C) x K; C aliases neither operand */
IAK3D_API int iak3d_mat_gemm_nt(iak3d_mat* C, double alpha, const iak3d_mat* A, const iak3d_mat* B, int32_t accumulate);
/* M := chol2inv(chol(M)) for a symmetric M; *status = IAK3D_NOT_SPD (M unchanged) where R's chol() stops */
IAK3D_API int iak3d_mat_spd_inverse(iak3d_mat* M, int32_t* status);
/* M (n x n) := setCIAK3D(...)$C of the dataset's supports (fitIAK3D.R:1005-1100); IAK3D_BAD_PARS like iak3d_cov_build */
IAK3D_API int iak3d_mat_cov_build(iak3d_ds* ds, const iak3d_pars* pars, iak3d_mat* M);
/* out[j] = sum_i A[i,j] * B[i,j]  ==  colSums(A * B) */
IAK3D_API int iak3d_mat_coldot(const iak3d_mat* A, const iak3d_mat* B, double* out);

/* ---- measurement helpers (bench.py) ---------------------------------------------------------- */
IAK3D_API int iak3d_sync(iak3d_ctx* ctx);
IAK3D_API int iak3d_timer_start(iak3d_ctx* ctx);                 /* cudaEventRecord on the context stream */
IAK3D_API int iak3d_timer_stop(iak3d_ctx* ctx, double* ms);      /* record + synchronize + elapsed */
IAK3D_API int iak3d_flush_l2(iak3d_ctx* ctx);                    /* writes a 256 MiB scratch buffer */
/* FP64 peak microbenchmarks: kind 0 = DFMA (vector), 1 = DMMA (mma.sync m8n8k4 f64), 2 = both at once (alternate warps):
 * a combined rate above either alone would mean two separate pipes -- measured: it is not. */
IAK3D_API int iak3d_fp64_peak(iak3d_ctx* ctx, int32_t kind, double* tflops);
/* per-stage device times (ms) of the last nll_terms_batch_enqueue: [0] table kernels, [1] covariance
 * build, [2] factorisation, [3] finish; and the algorithmic flop / byte counts of that launch. */
IAK3D_API int iak3d_last_stage_ms(iak3d_ctx* ctx, double* ms4, double* chol_flops, double* cov_bytes);

/* diagnostics: when the environment variable IAK3D_PROF is set, every tile task of the last nll_terms_batch records
 * device timestamps; out holds 11 int64 per task: problem, row block, column block, then 8 values -- start, K-loop
 * end, tile ready, diagonal/inverse ready, copy-out done, published, last dependency seen (ns), (smid << 32) | is_diag. */
IAK3D_API int iak3d_last_task_profile(iak3d_ctx* ctx, int64_t* out, int64_t cap_tasks, int64_t* ntasks);

#ifdef __cplusplus
}
#endif
#endif /* IAK3D_H */
