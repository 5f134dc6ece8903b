### iak3d_cuda.R -- drop-in GPU back end for the hot path of iak3d.
### source() this file AFTER the reference's own files (egEdgeroi.R:53-71) and after dyn.load("iak3dR.so")
### (r/r_shim.c).  It re-defines, with identical arguments and return values, the functions the optimiser and the
### predictor look up by name at call time:
###     setupIAK3D (adds a device handle), lndetANDinvCb, setCIAK3D, setCIAK3D2, nllIAK3D, gradnllIAK3D[.mc]
### and keeps everything else (readParsIAK3D, fitIAK3D, optimIt, makeXvX, ...) as the reference defines it.
### Forked workers cannot share a CUDA context: the glue forces the serial paths and batches those loops instead.

.iak3d_ref <- list(setupIAK3D = setupIAK3D, lndetANDinvCb = lndetANDinvCb, nllIAK3D = nllIAK3D, nllIAK3D_CL = nllIAK3D_CL,
                   setupIAK3D_CL = setupIAK3D_CL, makeXvX = makeXvX)
### An evaluation whose setupMats carries no device handle is an error, never a silent CPU evaluation.
.iak3d_state <- new.env()
.iak3d_state$ssre_warned <- FALSE ; .iak3d_state$cl <- NULL
.iak3d_modelx <- c(nugget = 0, spherical = 1, matern = 2, wendland = 3)
### every matrix handed to .Call is a double matrix (as.matrix() of integer-valued coordinates or depths is INTSXP; the shim reads REAL())
.iak3d_dbl <- function(x){ x <- as.matrix(x) ; storage.mode(x) <- 'double' ; x }

.iak3d_maxspl <- 12   # IAK3D_MAX_SPL
### the arguments of readParsIAK3D other than the vector, as the numeric vector r_shim.c's model_from() reads
.iak3d_model <- function(modelx, nud, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, prodSum, parBnds, lnTfmdData = FALSE){
  b <- function(nm){ v <- parBnds[[nm]] ; if(is.null(v)) c(NaN, NaN) else as.numeric(v)[1:2] }
  as.numeric(c(.iak3d_modelx[[modelx]], cmeOpt, as.integer(prodSum), as.integer(lnTfmdData), sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, nud,
               b('axBnds'), b('nuxBnds'), b('adBnds'), b('tau2Bnds'), b('sxd1Bnds'), if(is.null(parBnds$maxd)) NaN else parBnds$maxd))
}

.iak3d_pv <- function(parsBTfmd, modelx, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, quirk_scale = 1){
  tau <- function(v, t){ if(t == -1){ as.numeric(v)[1:2] }else{ c(0, 0) } }
  spl <- function(v, t){ out <- numeric(.iak3d_maxspl) ; if(t > 0){ out[seq(t)] <- as.numeric(v)[seq(t)] } ; out }
  nz <- function(v){ if(is.null(v) || length(v) == 0 || is.na(v)) NaN else as.numeric(v) }
  c(.iak3d_modelx[[modelx]], cmeOpt, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, 1,
    nz(parsBTfmd$ax), nz(parsBTfmd$nux), nz(parsBTfmd$ad), nz(parsBTfmd$nud),
    parsBTfmd$cx0, parsBTfmd$cx1, parsBTfmd$cd1, parsBTfmd$cxd0, parsBTfmd$cxd1, parsBTfmd$cme,
    tau(parsBTfmd$sdfdPars_cd1, sdfdType_cd1), tau(parsBTfmd$sdfdPars_cxd0, sdfdType_cxd0),
    tau(parsBTfmd$sdfdPars_cxd1, sdfdType_cxd1), quirk_scale,
    spl(parsBTfmd$sdfdPars_cd1, sdfdType_cd1), spl(parsBTfmd$sdfdPars_cxd0, sdfdType_cxd0), spl(parsBTfmd$sdfdPars_cxd1, sdfdType_cxd1),
    length(parsBTfmd$cssre), c(as.numeric(parsBTfmd$cssre), numeric(8))[1:8])   # site-specific random effects (fitIAK3D.R:1413)
}

### setupIAK3D: the id vectors live on the device.  Spline sd functions (sdfdType > 0): the interval-averaged spline
### bases on the unique depth intervals (iaCovMatern.R:520-539) are made here with the reference's own
### makeXnsClampedUG0 and handed to the library, which then evaluates setCApproxIAK3D[2] on the device.
.iak3d_spl_bases <- function(dI, types, sdfdKnots){
  nm <- c('cd1', 'cxd0', 'cxd1')
  lapply(seq(3), function(c){ if(types[c] > 0){
    makeXnsClampedUG0(x = dI, bdryKnots = sdfdKnots$bdryKnots, intKnots = sdfdKnots[[paste0('intKnots_', nm[c])]], paramVs = 1) }else{ NULL } })
}
setupIAK3D <- function(xData, dIData, nDscPts = 0, partSetup = FALSE, sdfdType_cd1 = 0, sdfdType_cxd0 = 0, sdfdType_cxd1 = 0,
                       sdfdKnots = NULL, siteIDData = NULL, XData = NULL, colnames4ssre = NULL){
  xData <- .iak3d_dbl(xData) ; dIData <- .iak3d_dbl(dIData)
  types <- c(sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1)
  h <- .Call('iak3dR_dataset', xData, dIData, NULL)
  siteLevels <- NULL
  if(!is.null(siteIDData)){   # site-specific random effects (iaCovMatern.R:440-460): the structures are never formed, see iak3d.h
    if(!all(XData[,1] == 1)){ stop('Error - the lmer-type mixed-effects option (if siteIDData given) only works if first column of X is 1s!') }
    siteLevels <- unique(as.character(siteIDData))
    .Call('iak3dR_set_ssre', h, as.integer(match(as.character(siteIDData), siteLevels) - 1L), .iak3d_dbl(XData[, colnames4ssre, drop = FALSE]))
  }
  if(max(types) > 0){
    dIU <- dIData[!duplicated(dIData), , drop = FALSE]                 # first-occurrence order == the library's ids
    B <- .iak3d_spl_bases(dIU, types, sdfdKnots)
    for(c in seq(3)){ if(!is.null(B[[c]])){ .Call('iak3dR_set_sdfd_spline', h, as.integer(c - 1), B[[c]]) } }
  }
  list(iak3d = h, xData = xData, dIData = dIData, n = nrow(xData), sdfdTypes = types, sdfdKnots = sdfdKnots,
       siteLevels = siteLevels, colnames4ssre = colnames4ssre, listStructs4ssre = if(is.null(siteIDData)) NULL else as.list(colnames4ssre),
       maxd = max(max(dIData), 2), Kx = TRUE) # 'Kx' non-NULL: the reference's is.null(setupMats$Kx) guards (fitIAK3D.R:965) stay quiet
}

lndetANDinvCb <- function(C, b = NULL){
  C <- .iak3d_dbl(C)
  if(nrow(C) < 256){ return(.iak3d_ref$lndetANDinvCb(C, b)) } # the p x p systems stay on the host
  tmp <- .Call('iak3dR_lndet_invCb', C, if(is.null(b)) NULL else .iak3d_dbl(b))
  list(lndetC = tmp[[1]], invCb = tmp[[2]])
}

setCIAK3D <- function(parsBTfmd, modelx, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, setupMats){
  tmp <- .Call('iak3dR_cov_build', setupMats$iak3d,
               .iak3d_pv(parsBTfmd, modelx, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt), TRUE)
  if(tmp[[1]] != 0){ return(list(C = NA, sigma2Vec = NA)) }
  list(C = forceSymmetric(Matrix(tmp[[2]])), sigma2Vec = tmp[[3]])
}

setCIAK3D2 <- function(parsBTfmd, modelx, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, setupMats){
  ### setupMats from setupIAK3D2(): set 1 = rows (prediction supports), set 2 = the data (a device dataset)
  pv <- .iak3d_pv(parsBTfmd, modelx, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt)
  if(max(sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1) > 0){               # setCApproxIAK3D2 (fitIAK3D.R:1121-1124)
    return(.Call('iak3dR_cov_cross_spl', setupMats$set2$iak3d, pv, .iak3d_dbl(setupMats$xData), .iak3d_dbl(setupMats$dIData), setupMats$Xspl1))
  }
  .Call('iak3dR_cov_cross', setupMats$set2$iak3d, pv, .iak3d_dbl(setupMats$xData), .iak3d_dbl(setupMats$dIData))
}
setupIAK3D2 <- function(xData, dIData, xData2, dIData2, sdfdType_cd1 = 0, sdfdType_cxd0 = 0, sdfdType_cxd1 = 0, sdfdKnots = NULL, ...){
  types <- c(sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1)
  list(xData = .iak3d_dbl(xData), dIData = .iak3d_dbl(dIData),
       set2 = setupIAK3D(xData2, dIData2, sdfdType_cd1 = sdfdType_cd1, sdfdType_cxd0 = sdfdType_cxd0, sdfdType_cxd1 = sdfdType_cxd1, sdfdKnots = sdfdKnots),
       Xspl1 = if(max(types) > 0) .iak3d_spl_bases(as.matrix(dIData), types, sdfdKnots) else list(NULL, NULL, NULL), Kx = TRUE)
}

### ---- lognormal data (lnTfmdData = TRUE): nrUpdatesIAK3DlnN on the Gram matrix of ONE bordered factorisation ----
### F = [X, z, s, V_ab (a <= b over iU)], s = sigma2Vec - diag(C) filled on the device; G = F' iC F carries every product the
### Newton iteration needs (u' iC v = cu' G cv, u' TK v = cu' GT cv); see iak3d_b200/api.py: lnN_design, nrUpdatesIAK3DlnN_gram.
.iak3d_lnN_design <- function(XData, zData, vXU, iU){
  n <- nrow(XData) ; pU <- length(iU)
  prs <- if(pU > 0) do.call(rbind, lapply(seq(pU), function(a){ cbind(a, a:pU) })) else matrix(0, 0, 2)
  V <- if(pU > 0) sapply(seq(nrow(prs)), function(k){ vXU[(seq(n) - 1) * pU + prs[k, 1], prs[k, 2]] }) else NULL
  list(F = cbind(XData, zData, 0, V), pairs = prs)
}
.iak3d_nr_lnN <- function(G, lndetC, n, p, iU, prs){
  G <- 0.5 * (G + t(G)) ; q <- nrow(G) ; pU <- length(iU) ; iK <- setdiff(seq(p), iU) ; zc <- p + 1 ; sc <- p + 2
  e <- function(j){ v <- numeric(q) ; v[j] <- 1 ; v }
  vcol <- function(a, b){ p + 2 + which(prs[, 1] == min(a, b) & prs[, 2] == max(a, b)) }
  XiCX <- G[1:p, 1:p, drop = FALSE]
  if(pU == 0){
    cc <- e(zc) - 0.5 * e(sc) ; XiCz <- G[1:p, , drop = FALSE] %*% cc ; b <- solve(XiCX, XiCz)
    res <- as.numeric(t(cc) %*% G %*% cc - 2 * t(b) %*% XiCz + t(b) %*% XiCX %*% b)
    return(list(betahat = b, nll = 0.5 * (n * log(2 * pi) + lndetC + res), fim = -XiCX, noits = 0, errFlag = 0))
  }
  betaU <- solve(XiCX, G[1:p, zc])[iU]
  if(length(iK) > 0){ GKK <- G[iK, iK, drop = FALSE] ; GT <- G - G[, iK, drop = FALSE] %*% solve(GKK, G[iK, , drop = FALSE]) }else{ GT <- G }
  gh <- function(bU){
    cres <- e(zc) - 0.5 * e(sc) ; Ca <- matrix(0, q, pU)
    for(a in seq(pU)){
      cres <- cres - bU[a] * e(iU[a]) ; ca <- e(iU[a])
      for(b in seq(pU)){ ca <- ca + bU[b] * e(vcol(a, b)) ; cres <- cres - 0.5 * bU[a] * bU[b] * e(vcol(a, b)) }
      Ca[, a] <- ca
    }
    Tres <- GT %*% cres
    fim <- t(Ca) %*% GT %*% Ca
    vT <- matrix(sapply(seq(pU), function(b){ sapply(seq(pU), function(a){ Tres[vcol(a, b)] }) }), pU, pU)
    hess <- -vT + fim ; betahat <- matrix(NA, p, 1) ; betahat[iU] <- bU
    if(length(iK) > 0){ betahat[iK] <- solve(GKK, G[iK, , drop = FALSE] %*% cres) }
    list(nll = as.numeric(0.5 * (n * log(2 * pi) + lndetC + t(cres) %*% Tres)), grad = -t(Ca) %*% Tres,
         hess = 0.5 * (hess + t(hess)), fim = 0.5 * (fim + t(fim)), betahat = betahat)
  }
  cur <- gh(betaU) ; tolNr <- 1E-6 ; noits <- 0 ; nllPrev <- Inf
  while((noits < 20) & ((cur$nll < (nllPrev - tolNr)) | (noits < 4))){
    nllPrev <- cur$nll
    ### (nrUpdatesIAK3DlnN.R:109 tests `checkHess$cholC` on the value of solve(), which cannot run; intended logic:)
    step <- if(noits >= 2) try(solve(cur$hess, cur$grad), silent = TRUE) else NULL
    if(is.null(step) || is.character(step)){ step <- try(solve(cur$fim, cur$grad), silent = TRUE) }
    if(is.character(step)){ return(c(cur, list(noits = noits, errFlag = 1))) }
    betaU <- betaU - as.numeric(step) ; cur <- gh(betaU) ; noits <- noits + 1
  }
  cur$nll <- round(cur$nll / (tolNr * 10)) * (tolNr * 10)
  c(cur, list(noits = noits, errFlag = if(noits == 20) 2 else 0))
}
.iak3d_nll_lnN <- function(pars, zData, XData, vXU, iU, modelx, nud, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, prodSum,
                           setupMats, parBnds, rtnAll, attachBigMats){
  XData <- as.matrix(XData) ; n <- length(zData) ; p <- ncol(XData)
  parsBTfmd <- readParsIAK3D(pars = pars, modelx = modelx, nud = nud, sdfdType_cd1 = sdfdType_cd1, sdfdType_cxd0 = sdfdType_cxd0,
                             sdfdType_cxd1 = sdfdType_cxd1, cmeOpt = cmeOpt, prodSum = prodSum, parBnds = parBnds, lnTfmdData = TRUE)
  pv <- .iak3d_pv(parsBTfmd, modelx, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt)
  des <- .iak3d_lnN_design(XData, zData, vXU, iU) ; q <- ncol(des$F)
  .Call('iak3dR_set_W', setupMats$iak3d, des$F) ; .Call('iak3dR_set_lnN_column', setupMats$iak3d, as.integer(p + 1))
  tmp <- .Call('iak3dR_nll_terms', setupMats$iak3d, matrix(pv, ncol = 1))
  .Call('iak3dR_set_lnN_column', setupMats$iak3d, as.integer(-1))
  if(tmp[[1]][1] != 0){ if(rtnAll){ return(list(nll = 9E99, pars = pars, parsBTfmd = parsBTfmd, betahat = NA, vbetahat = NA, cxdhat = NA)) }else{ return(9E99) } }
  G <- matrix(tmp[[3]][, 1], q, q)
  nr <- .iak3d_nr_lnN(G, tmp[[2]][1], n, p, iU, des$pairs)
  nll <- if(is.na(nr$nll) || is.infinite(nr$nll)) 9E99 else nr$nll
  if(!rtnAll){ return(nll) }
  ### vbetahat = solve(fim) with every column treated as varying and TK = iC (fitIAK3D.R:768-773)
  Gs <- 0.5 * (G + t(G)) ; Ca <- diag(q)[, 1:p, drop = FALSE]
  vcol <- function(a, b){ p + 2 + which(des$pairs[, 1] == min(a, b) & des$pairs[, 2] == max(a, b)) }
  for(a in seq_along(iU)){ for(b in seq_along(iU)){ Ca[vcol(a, b), iU[a]] <- Ca[vcol(a, b), iU[a]] + nr$betahat[iU[b]] } }
  vbetahat <- solve(t(Ca) %*% Gs %*% Ca)
  s2 <- .Call('iak3dR_cov_build', setupMats$iak3d, pv, FALSE)[[3]] ; dC <- .Call('iak3dR_cov_diag', setupMats$iak3d, pv)[[2]]
  muhat <- setmuIAK3D(XData = XData, vXU = vXU, iU = iU, beta = nr$betahat, diagC = dC, sigma2Vec = s2)
  out <- list(nll = nll, pars = pars, parsBTfmd = parsBTfmd, betahat = nr$betahat, vbetahat = vbetahat, cxdhat = 1, sigma2Vec = s2,
              XData = XData, vXU = vXU, iU = iU, diagC = dC)
  .Call('iak3dR_set_W', setupMats$iak3d, cbind(XData, zData))
  if(attachBigMats){
    fit <- .Call('iak3dR_fit_res', setupMats$iak3d, pv, as.numeric(nr$betahat), as.matrix(vbetahat), as.numeric(zData - muhat))
    sm <- .Call('iak3dR_fit_small', fit, as.integer(n), as.integer(p))
    out$iak3dFit <- fit ; out$iCX <- sm[[1]] ; out$iCz_muhat <- sm[[2]]
  }
  out
}

### the REML / ML tail of nllIAK3D (fitIAK3D.R:821-880) on the host: O(p^3)
.iak3d_tail <- function(WiAW, lndetA, n, p, useReml){
  XiAX <- WiAW[1:p, 1:p, drop = FALSE] ; XiAz <- WiAW[1:p, p+1, drop = FALSE] ; ziAz <- WiAW[p+1, p+1]
  tmp <- .iak3d_ref$lndetANDinvCb(XiAX, XiAz)
  if(is.na(tmp$lndetC) || is.infinite(tmp$lndetC)){ return(list(nll = 9E99)) }
  betahat <- matrix(tmp$invCb, ncol = 1)
  resiAres <- as.numeric(ziAz - 2 * t(betahat) %*% XiAz + t(betahat) %*% XiAX %*% betahat)
  cxdhat <- if(useReml) resiAres / (n - p) else resiAres / n
  lndetC <- lndetA + n * log(cxdhat) ; lndetXiCX <- tmp$lndetC - p * log(cxdhat) ; resiCres <- resiAres / cxdhat
  nll <- if(useReml) 0.5 * ((n - p) * log(2 * pi) + lndetC + lndetXiCX + resiCres) else 0.5 * (n * log(2 * pi) + lndetC + resiCres)
  if(is.na(nll) || is.nan(nll) || is.infinite(nll)){ nll <- 9E99 }
  list(nll = as.numeric(nll), betahat = betahat, cxdhat = cxdhat, XiAX = XiAX)
}

nllIAK3D <- function(pars, zData, XData, vXU, iU, modelx, nud, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, prodSum,
                     setupMats = NULL, parBnds, useReml, lnTfmdData, rtnAll = F, forCompLik = FALSE, attachBigMats = FALSE, nCores = 8){
  if(is.null(setupMats$iak3d)){
    stop('iak3d: setupMats carries no device dataset -- build it with the setupIAK3D of r/iak3d_cuda.R (there is no silent CPU fallback)')
  }
  ### composite likelihood: the terms of this unit were computed with all the others in one batched launch (nllIAK3D_CL below)
  if(forCompLik && !(rtnAll && attachBigMats) && !is.null(.iak3d_state$cl) && !is.null(setupMats$iak3d_unit)){
    u <- .iak3d_state$cl$res ; k <- match(setupMats$iak3d_unit, .iak3d_state$cl$units) ; p <- ncol(as.matrix(XData))
    if(!is.na(k)){
      if(u[[1]][k] != 0){ return(list(pars = pars, parsBTfmd = .iak3d_state$cl$parsBTfmd, sigma2Vec = NA, WiAW = matrix(NA, p+1, p+1), lndetA = NA)) }
      return(list(pars = pars, parsBTfmd = .iak3d_state$cl$parsBTfmd, sigma2Vec = NA,
                  WiAW = as.matrix(forceSymmetric(Matrix(matrix(u[[3]][, k], p + 1, p + 1)))), lndetA = u[[2]][k]))
    }
  }
  if(lnTfmdData){
    return(.iak3d_nll_lnN(pars, zData, XData, vXU, iU, modelx, nud, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, prodSum,
                          setupMats, parBnds, rtnAll, attachBigMats))
  }
  XData <- .iak3d_dbl(XData) ; zData <- as.numeric(zData) ; n <- length(zData) ; p <- ncol(XData)
  if(!rtnAll && !forCompLik){                      # what optim() calls: the vector goes to the library as it is (one .Call)
    .Call('iak3dR_set_W', setupMats$iak3d, cbind(XData, zData))
    return(.Call('iak3dR_nll_vector', setupMats$iak3d,
                 .iak3d_model(modelx, nud, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, prodSum, parBnds), as.numeric(pars), as.logical(useReml)))
  }
  parsBTfmd <- readParsIAK3D(pars = pars, modelx = modelx, nud = nud, sdfdType_cd1 = sdfdType_cd1, sdfdType_cxd0 = sdfdType_cxd0,
                             sdfdType_cxd1 = sdfdType_cxd1, cmeOpt = cmeOpt, prodSum = prodSum, parBnds = parBnds, lnTfmdData = lnTfmdData)
  .Call('iak3dR_set_W', setupMats$iak3d, cbind(XData, zData))
  pv <- .iak3d_pv(parsBTfmd, modelx, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt)
  if(forCompLik || (rtnAll && !attachBigMats)){
    tmp <- .Call('iak3dR_nll_terms_iAW', setupMats$iak3d, pv)
    status <- tmp[[1]] ; lndetA <- tmp[[2]] ; WiAW <- tmp[[3]] ; iAW <- tmp[[4]]
  }else{
    tmp <- .Call('iak3dR_nll_terms', setupMats$iak3d, matrix(pv, ncol = 1))
    status <- tmp[[1]][1] ; lndetA <- tmp[[2]][1] ; WiAW <- matrix(tmp[[3]][, 1], p + 1, p + 1) ; iAW <- NULL
  }
  if(status != 0){
    if(forCompLik){ return(list(pars = pars, parsBTfmd = parsBTfmd, sigma2Vec = NA, WiAW = matrix(NA, p+1, p+1), lndetA = NA)) }
    if(rtnAll){ return(list(nll = 9E99, pars = pars, parsBTfmd = parsBTfmd, betahat = NA, vbetahat = NA, cxdhat = NA)) }
    return(9E99)
  }
  WiAW <- as.matrix(forceSymmetric(Matrix(WiAW)))
  if(forCompLik){ return(list(pars = pars, parsBTfmd = parsBTfmd, sigma2Vec = NA, WiAW = WiAW, lndetA = lndetA, iAW = iAW)) }
  tl <- .iak3d_tail(WiAW, lndetA, n, p, useReml)
  if(!rtnAll){ return(tl$nll) }
  cxdhat <- tl$cxdhat
  for(k in c('cx0', 'cx1', 'cd1', 'cxd0', 'cxd1')){ parsBTfmd[[k]] <- cxdhat * parsBTfmd[[k]] }
  if(cmeOpt == 1){ parsBTfmd$cme <- cxdhat * parsBTfmd$cme }
  if(!is.null(parsBTfmd$cssre)){ parsBTfmd$cssre <- cxdhat * parsBTfmd$cssre }           # fitIAK3D.R:863
  vbetahat <- cxdhat * solve(tl$XiAX)
  out <- list(nll = tl$nll, pars = pars, parsBTfmd = parsBTfmd, betahat = tl$betahat, vbetahat = vbetahat, cxdhat = cxdhat,
              XData = XData, vXU = vXU, iU = iU)
  if(attachBigMats){
    ### the big matrices stay on the device as a kept factorisation; predictIAK3D uses lmmFit$iak3dFit
    pvS <- .iak3d_pv(parsBTfmd, modelx, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, quirk_scale = cxdhat)
    fit <- .Call('iak3dR_fit', setupMats$iak3d, pvS, as.numeric(tl$betahat), as.matrix(vbetahat))
    sm <- .Call('iak3dR_fit_small', fit, as.integer(n), as.integer(p))
    out$iak3dFit <- fit ; out$iCX <- sm[[1]] ; out$iCz_muhat <- sm[[2]]
  }else{
    out$iAX <- iAW[, 1:p, drop = FALSE] ; out$iAz <- iAW[, p+1, drop = FALSE]
  }
  out
}

### forward-difference gradient: npar + 1 evaluations in ONE device batch (fitIAK3D.R:1865-1886; the .mc variant
### forks and cannot be used with CUDA)
gradnllIAK3D <- function(pars, zData, XData, vXU, iU, modelx, nud, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, prodSum,
                         setupMats, parBnds, useReml, lnTfmdData, ...){
  XData <- .iak3d_dbl(XData) ; zData <- as.numeric(zData) ; n <- length(zData) ; p <- ncol(XData) ; delta <- 1E-6
  .Call('iak3dR_set_W', setupMats$iak3d, cbind(XData, zData))
  if(!lnTfmdData){                                 # transform, npar + 1 evaluations in one launch, tails, differences: one .Call
    return(.Call('iak3dR_gradnll_fd', setupMats$iak3d,
                 .iak3d_model(modelx, nud, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, prodSum, parBnds), as.numeric(pars), delta, as.logical(useReml))[[1]])
  }
  cand <- cbind(pars, pars + delta * diag(length(pars)))
  pm <- apply(cand, 2, function(v){
    .iak3d_pv(readParsIAK3D(pars = v, modelx = modelx, nud = nud, sdfdType_cd1 = sdfdType_cd1, sdfdType_cxd0 = sdfdType_cxd0,
                            sdfdType_cxd1 = sdfdType_cxd1, cmeOpt = cmeOpt, prodSum = prodSum, parBnds = parBnds, lnTfmdData = lnTfmdData),
              modelx, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt) })
  tmp <- .Call('iak3dR_nll_terms', setupMats$iak3d, pm)
  f <- sapply(seq(ncol(cand)), function(i){
    if(tmp[[1]][i] != 0){ return(9E99) }
    .iak3d_tail(as.matrix(forceSymmetric(Matrix(matrix(tmp[[3]][, i], p+1, p+1)))), tmp[[2]][i], n, p, useReml)$nll })
  g <- (f[-1] - f[1]) / delta
  g[is.na(g)] <- 0
  g
}
gradnllIAK3D.mc <- gradnllIAK3D

### the same gradient by ONE factorisation + explicit inverse + one pass over it (iak3d_nll_grad; not in the reference).
### Faster than the batch from n ~ 4000 on (B200: 1.7x at 5000, 2.8x at 20000); pass as `gr` to optimIt in place of gradnllIAK3D.
gradnllIAK3D_analytic <- function(pars, zData, XData, vXU, iU, modelx, nud, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, prodSum,
                                  setupMats, parBnds, useReml, lnTfmdData, ...){
  XData <- .iak3d_dbl(XData) ; zData <- as.numeric(zData) ; delta <- 1E-6
  .Call('iak3dR_set_W', setupMats$iak3d, cbind(XData, zData))
  cand <- cbind(pars, pars + delta * diag(length(pars)))
  pm <- apply(cand, 2, function(v){
    .iak3d_pv(readParsIAK3D(pars = v, modelx = modelx, nud = nud, sdfdType_cd1 = sdfdType_cd1, sdfdType_cxd0 = sdfdType_cxd0,
                            sdfdType_cxd1 = sdfdType_cxd1, cmeOpt = cmeOpt, prodSum = prodSum, parBnds = parBnds, lnTfmdData = lnTfmdData),
              modelx, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt) })
  tmp <- .Call('iak3dR_nll_grad', setupMats$iak3d, pm, rep(delta, length(pars)), useReml)
  if(tmp[[1]] != 0){ return(numeric(length(pars))) }
  g <- 0.5 * tmp[[4]]
  g[is.na(g)] <- 0
  g
}


### ---- the per-batch body of predictIAK3D / profilePredictIAK3D (predictIAK3D.R:179-255, 440-491) ----
### call instead of lines 179-208 + 229-255 when lmmFit$iak3dFit is present; returns list(z, v) before the back-transform
iak3dPredictBatch <- function(lmmFit, xMapThis, dIMapThis, XMapThis, vXUMapThis = NULL){
  types <- c(lmmFit$sdfdType_cd1, lmmFit$sdfdType_cxd0, lmmFit$sdfdType_cxd1)
  spl <- if(max(types) > 0) .iak3d_spl_bases(as.matrix(dIMapThis), types, lmmFit$sdfdKnots) else NULL
  tmp <- .Call('iak3dR_predict_terms', lmmFit$iak3dFit, .iak3d_dbl(xMapThis), .iak3d_dbl(dIMapThis), .iak3d_dbl(XMapThis), spl, FALSE)
  if(!lmmFit$lnTfmdData){ return(list(z = tmp[[1]], v = tmp[[2]])) }
  ### lognormal data: muhat from setmuIAK3D, no beta uncertainty in the variance (predictIAK3D.R:232, 251)
  mu <- setmuIAK3D(XData = XMapThis, vXU = vXUMapThis, iU = lmmFit$iU, beta = lmmFit$betahat, diagC = tmp[[3]], sigma2Vec = tmp[[4]])
  list(z = as.numeric(mu) + tmp[[5]], v = tmp[[3]] - tmp[[6]])
}

### gradHessIAK3D (nrUpdatesIAK3D.R:118-216) on the device; nrUpdatesIAK3D's loop (:1-116) runs unchanged around it
gradHessIAK3D_gpu <- function(setupMats, parsBTfmd, modelx, sdfdTypes, lnvPars){
  pv <- .iak3d_pv(parsBTfmd, modelx, sdfdTypes[1], sdfdTypes[2], sdfdTypes[3], if(length(lnvPars) == 6) 1 else 0) ; pv[6] <- 0
  tmp <- .Call('iak3dR_gradhess', setupMats$iak3d, pv, as.numeric(lnvPars))
  if(tmp[[1]] != 0){ return(list(nll = NA)) }
  list(nll = tmp[[2]], grad = tmp[[3]], hess = tmp[[4]], fim = tmp[[5]], betahat = tmp[[6]], vbetahat = tmp[[7]])
}


### ---- predMatsIAK3D_CLEV_ByBlock (compositeLikIAK3D.R:609-798) with every large matrix kept on the device ----
### Thin R handles over iak3d_mat_*: m$new / m$sub / m$mm are `matrix(0, .)`, `M[rows, cols]` and `A %*% t(B)` of the
### reference; chol2inv(chol(.)) and solve(.) are iak3dR_mat_spd_inverse.  Same arguments and return value as the reference.
.iak3d_m <- list(
  new = function(nr, nc){ .Call('iak3dR_mat', nr, nc, NULL) },
  from = function(A){ .Call('iak3dR_mat', 0, 0, as.matrix(A)) },
  get = function(M, r0, c0, nr, nc){ .Call('iak3dR_mat_get', M, as.numeric(c(r0, c0, nr, nc))) },
  copy = function(D, S, dr0 = 0, dc0 = 0, sr0 = 0, sc0 = 0, nr, nc, tr = 0, alpha = 1, beta = 0){
    .Call('iak3dR_mat_copy', D, S, as.numeric(c(dr0, dc0, sr0, sc0, nr, nc, tr, alpha, beta))) ; D },
  sub = function(S, r0, c0, nr, nc, tr = 0, alpha = 1){
    D <- if(tr == 1){ .Call('iak3dR_mat', nc, nr, NULL) }else{ .Call('iak3dR_mat', nr, nc, NULL) }
    if(tr == 1){ .iak3d_m$copy(D, S, sr0 = r0, sc0 = c0, nr = nc, nc = nr, tr = 1, alpha = alpha) }else{
      .iak3d_m$copy(D, S, sr0 = r0, sc0 = c0, nr = nr, nc = nc, alpha = alpha) } },
  gather = function(S, rr, cc){   # S[c(rows...), c(cols...)] for lists of c(start0, length)
    D <- .Call('iak3dR_mat', sum(sapply(rr, `[`, 2)), sum(sapply(cc, `[`, 2)), NULL) ; ro <- 0
    for(r in rr){ co <- 0
      for(k in cc){ if(r[2] > 0 && k[2] > 0){ .iak3d_m$copy(D, S, ro, co, r[1], k[1], r[2], k[2]) } ; co <- co + k[2] }
      ro <- ro + r[2] }
    D },
  sym = function(S, n){ D <- .Call('iak3dR_mat', n, n, NULL)
    .iak3d_m$copy(D, S, nr = n, nc = n, alpha = 0.5) ; .iak3d_m$copy(D, S, nr = n, nc = n, tr = 1, alpha = 0.5, beta = 1) },
  mm = function(A, B, nr, nc, alpha = 1){ D <- .Call('iak3dR_mat', nr, nc, NULL) ; .Call('iak3dR_mat_gemm_nt', D, alpha, A, B, 0L) ; D },
  acc = function(D, A, B, alpha = 1){ .Call('iak3dR_mat_gemm_nt', D, alpha, A, B, 1L) ; D },
  inv = function(M){ .Call('iak3dR_mat_spd_inverse', M) ; M })

predMatsIAK3D_CLEV_ByBlock <- function(z_muhat , XData , xMap , dIMap , iData = seq(length(z_muhat)) , lmmFit){
  m_ <- .iak3d_m
  blockMap <- getBlocksFromLocations(xMap = xMap , compLikMats = lmmFit$compLikMats)
  zk <- vk <- matrix(NA , length(blockMap) , 1)
  pv <- .iak3d_pv(lmmFit$parsBTfmd, lmmFit$modelx, lmmFit$sdfdType_cd1, lmmFit$sdfdType_cxd0, lmmFit$sdfdType_cxd1, lmmFit$cmeOpt)
  for (iBlock in 1:max(blockMap)){
    iMapThis <- which(blockMap == iBlock) ; m <- length(iMapThis)
    if(m == 0){ next }
    iBlocksAdj <- getAdjacentBlocks(iBlock = iBlock , compLikMats = lmmFit$compLikMats)
    iDatak <- intersect(lmmFit$compLikMats$listBlocks[[iBlock]]$i , iData) ; nk <- length(iDatak)
    iDatalList <- lapply(iBlocksAdj, function(l){ intersect(lmmFit$compLikMats$listBlocks[[l]]$i , iData) })
    nl <- sapply(iDatalList, length) ; s2 <- m + nk + cumsum(c(0, nl))[seq_along(nl)]          # zero-based starts of the i2 ranges
    iThis <- c(iDatak , unlist(iDatalList))
    sm <- setupIAK3D(xData = rbind(xMap[iMapThis,,drop=FALSE] , lmmFit$xData[iThis,,drop=FALSE]) ,
                     dIData = rbind(dIMap[iMapThis,,drop=FALSE] , as.matrix(lmmFit$dIData[iThis,,drop=FALSE])) , nDscPts = 0 ,
                     sdfdType_cd1 = lmmFit$sdfdType_cd1 , sdfdType_cxd0 = lmmFit$sdfdType_cxd0 , sdfdType_cxd1 = lmmFit$sdfdType_cxd1 ,
                     sdfdKnots = lmmFit$sdfdKnots)
    Call <- .Call('iak3dR_mat_cov_build', sm$iak3d, pv)                                          # C0klAll, :671
    C0k <- m_$sub(Call, 0, 0, m + nk, m + nk) ; C00 <- m_$sub(Call, 0, 0, m, m)
    A0 <- m_$new(m, m) ; b0 <- m_$new(m, 1) ; Bk0 <- m_$new(m, m + nk) ; QL <- list()
    for(i in seq_along(iBlocksAdj)){
      kl <- list(c(m, nk), c(s2[i], nl[i])) ; nkl <- nk + nl[i]
      iS <- m_$inv(m_$gather(Call, kl, kl))                                                      # :717
      C0kl <- m_$gather(Call, list(c(0, m)), kl)
      Tt <- m_$mm(C0kl, iS, m, nkl)
      IF <- m_$acc(m_$copy(m_$new(m, m), C00, nr = m, nc = m), C0kl, Tt, alpha = -1)             # :720
      iIF <- m_$inv(m_$sym(IF, m))                                                               # :722-725
      Q12 <- m_$mm(iIF, m_$sub(Tt, 0, 0, m, nkl, tr = 1), m, nkl, alpha = -1)                    # :726
      QL[[i]] <- m_$sub(Q12, 0, nk, m, nl[i])
      m_$copy(A0, iIF, nr = m, nc = m, beta = 1)
      m_$acc(b0, Q12, m_$from(matrix(z_muhat[c(iDatak , iDatalList[[i]])], nrow = 1)), alpha = -1)
      m_$copy(Bk0, iIF, nr = m, nc = m, beta = 1) ; m_$copy(Bk0, Q12, dc0 = m, nr = m, nc = nk, beta = 1)
    }
    if(lmmFit$compLikMats$compLikOptn == 3){                                                     # :748-770
      nnon <- length(lmmFit$compLikMats$listBlocks) - length(iBlocksAdj) - 1
      iCkk <- m_$inv(m_$sub(Call, m, m, nk, nk)) ; C0_k <- m_$sub(Call, 0, m, m, nk)
      C0kiC <- m_$mm(C0_k, iCkk, m, nk)
      iIF <- m_$inv(m_$sym(m_$acc(m_$copy(m_$new(m, m), C00, nr = m, nc = m), C0kiC, C0_k, alpha = -1), m))
      iIFC <- m_$mm(iIF, m_$sub(C0kiC, 0, 0, m, nk, tr = 1), m, nk)
      m_$copy(A0, iIF, nr = m, nc = m, alpha = nnon, beta = 1)
      m_$acc(b0, iIFC, m_$from(matrix(z_muhat[iDatak], nrow = 1)), alpha = nnon)
      m_$copy(Bk0, iIF, nr = m, nc = m, alpha = nnon, beta = 1) ; m_$copy(Bk0, iIFC, dc0 = m, nr = m, nc = nk, alpha = -nnon, beta = 1)
    }
    J0 <- m_$mm(m_$mm(Bk0, C0k, m, m + nk), Bk0, m, m)                                           # :772
    for(i in seq_along(iBlocksAdj)){
      m_$acc(J0, m_$mm(Bk0, m_$sub(Call, s2[i], 0, nl[i], m + nk), m, nl[i]), QL[[i]], alpha = 2)
      for(j in seq_along(iBlocksAdj)){
        m_$acc(J0, m_$mm(QL[[i]], m_$sub(Call, s2[j], s2[i], nl[j], nl[i]), m, nl[j]), QL[[j]]) }
    }
    iA0 <- m_$inv(m_$sym(A0, m)) ; J0s <- m_$sym(J0, m)                                          # :784-788
    zk[iMapThis,1] <- m_$get(m_$mm(iA0, m_$sub(b0, 0, 0, m, 1, tr = 1), m, 1), 0, 0, m, 1)
    vk[iMapThis,1] <- .Call('iak3dR_mat_coldot', iA0, m_$mm(J0s, iA0, m, m))
    rm(Call, C0k, C00, A0, b0, Bk0, QL, J0, iA0, J0s) ; gc(FALSE)       # finalizers release the device matrices of this block
  }
  .Call('iak3dR_mat_trim')
  return(list('zk' = zk , 'vk' = vk))
}


### ---- composite likelihood (compositeLikIAK3D.R:6-307) ----
### The reference evaluates nllIAK3D(forCompLik = TRUE) once per block pair / block, forked with mcmapply (:54) or in a loop
### (:86-104): here ALL units of one evaluation go through ONE batched launch (iak3dR_nll_terms_units; no A^-1 W
### back-substitution), and the reference's own nllIAK3D_CL then runs serially over cached per-unit results, so its tail
### (:181-307, compLikOptn 1-4) is untouched.  setupIAK3D_CL tags every unit's setupMats with its index.
setupIAK3D_CL <- function(xData , dIData , nDscPts = 0 , partSetup = FALSE , compLikMats = NULL , sdfdType_cd1 = 0 , sdfdType_cxd0 = 0 ,
                          sdfdType_cxd1 = 0 , sdfdKnots = NULL){
  sm <- .iak3d_ref$setupIAK3D_CL(xData, dIData, nDscPts, partSetup, compLikMats, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, sdfdKnots)
  for(i in seq_along(sm)){ if(!is.null(sm[[i]])){ sm[[i]]$iak3d_unit <- i } }
  sm
}
nllIAK3D_CL <- function(pars , zData , XData , modelx , nud , sdfdType_cd1 , sdfdType_cxd0 , sdfdType_cxd1 , cmeOpt , prodSum , setupMats = NULL ,
                        parBnds , useReml , compLikMats , parallel_nll_4cl = TRUE , rtnAll = F , attachBigMats = rtnAll , nCores = 1){
  XData <- as.matrix(XData)
  nP <- nrow(compLikMats$subsetPairsAdj) ; nB <- length(compLikMats$listBlocks) ; optn <- compLikMats$compLikOptn
  rowsOf <- function(i){ if(i <= nP){ c(compLikMats$listBlocks[[compLikMats$subsetPairsAdj[i,1]]]$i , compLikMats$listBlocks[[compLikMats$subsetPairsAdj[i,2]]]$i)
                         }else{ compLikMats$listBlocks[[i - nP]]$i } }
  units <- c(if(optn < 4) seq(nP) else NULL , if(optn %in% c(1, 3, 4)) nP + seq(nB) else NULL)
  units <- units[sapply(units, function(i){ !is.null(setupMats[[i]]$iak3d) })]
  if(length(units) > 0 && !(rtnAll && attachBigMats)){
    parsBTfmd <- readParsIAK3D(pars = pars, modelx = modelx, nud = nud, sdfdType_cd1 = sdfdType_cd1, sdfdType_cxd0 = sdfdType_cxd0,
                               sdfdType_cxd1 = sdfdType_cxd1, cmeOpt = cmeOpt, prodSum = prodSum, parBnds = parBnds, lnTfmdData = FALSE)
    for(i in units){ r <- rowsOf(i) ; .Call('iak3dR_set_W', setupMats[[i]]$iak3d, cbind(XData[r,,drop=FALSE], zData[r])) }
    res <- .Call('iak3dR_nll_terms_units', lapply(units, function(i){ setupMats[[i]]$iak3d }),
                 .iak3d_pv(parsBTfmd, modelx, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt))
    .iak3d_state$cl <- list(units = units, res = res, parsBTfmd = parsBTfmd)
    on.exit(.iak3d_state$cl <- NULL)
  }
  .iak3d_ref$nllIAK3D_CL(pars = pars, zData = zData, XData = XData, modelx = modelx, nud = nud, sdfdType_cd1 = sdfdType_cd1,
                         sdfdType_cxd0 = sdfdType_cxd0, sdfdType_cxd1 = sdfdType_cxd1, cmeOpt = cmeOpt, prodSum = prodSum, setupMats = setupMats,
                         parBnds = parBnds, useReml = useReml, compLikMats = compLikMats, parallel_nll_4cl = FALSE, rtnAll = rtnAll,
                         attachBigMats = attachBigMats, nCores = 1)
}

### predMatsIAK3D_CL_ByBlock (compositeLikIAK3D.R:471-607): for the map supports of block k, the kriging terms against the data
### of every adjacent pair (k, l), averaged over the neighbours.  Each pair is one kept factorisation on the device;
### IAK3D_FIT_CROSS_SQUARE reproduces the reference's Ckh = block of setCIAK3D of the joined supports (:540-552).
predMatsIAK3D_CL_ByBlock <- function(z_muhat , XData , xMap , dIMap , iData = seq(length(z_muhat)) , lmmFit){
  XData <- as.matrix(XData) ; xMap <- as.matrix(xMap) ; dIMap <- as.matrix(dIMap)
  blockMap <- getBlocksFromLocations(xMap = xMap , compLikMats = lmmFit$compLikMats)
  nxMap <- length(blockMap) ; n <- length(z_muhat) ; p <- ncol(XData)
  diagCkk <- diagCkhiCChk <- CkhiCz_muhat <- matrix(NA , nxMap , 1) ; CkhiCX <- matrix(NA , nxMap , p)
  types <- c(lmmFit$sdfdType_cd1, lmmFit$sdfdType_cxd0, lmmFit$sdfdType_cxd1)
  pv <- .iak3d_pv(lmmFit$parsBTfmd, lmmFit$modelx, types[1], types[2], types[3], lmmFit$cmeOpt)
  for (iBlock in 1:max(blockMap)){
    iMapThis <- which(blockMap == iBlock) ; m <- length(iMapThis)
    if(m == 0){ next }
    iBlocksAdj <- getAdjacentBlocks(iBlock = iBlock , compLikMats = lmmFit$compLikMats)
    iDatak <- lmmFit$compLikMats$listBlocks[[iBlock]]$i
    if(length(iData) < n){ iDatak <- intersect(iDatak , iData) }
    spl <- if(max(types) > 0) .iak3d_spl_bases(dIMap[iMapThis,,drop=FALSE], types, lmmFit$sdfdKnots) else NULL
    qS <- czS <- numeric(m) ; cxS <- matrix(0 , m , p)
    for(i in seq_along(iBlocksAdj)){
      iDatal <- lmmFit$compLikMats$listBlocks[[iBlocksAdj[i]]]$i
      if(length(iData) < n){ iDatal <- intersect(iDatal , iData) }
      iThis <- c(iDatak , iDatal)
      sm <- setupIAK3D(xData = lmmFit$xData[iThis,,drop=FALSE] , dIData = as.matrix(lmmFit$dIData[iThis,,drop=FALSE]) , sdfdType_cd1 = types[1] ,
                       sdfdType_cxd0 = types[2] , sdfdType_cxd1 = types[3] , sdfdKnots = lmmFit$sdfdKnots)
      .Call('iak3dR_set_W', sm$iak3d, cbind(XData[iThis,,drop=FALSE], z_muhat[iThis]))
      fit <- .Call('iak3dR_fit_res', sm$iak3d, pv, numeric(p), diag(p), as.numeric(z_muhat[iThis]))
      .Call('iak3dR_fit_option', fit, 1L, 1L) ; .Call('iak3dR_fit_option', fit, 2L, 1L)     # keep Ckh iC X; Ckh as a block of the square C
      tmp <- .Call('iak3dR_predict_terms', fit, xMap[iMapThis,,drop=FALSE], dIMap[iMapThis,,drop=FALSE], matrix(0, m, p), spl, TRUE)
      if(i == 1){ diagCkk[iMapThis,1] <- tmp[[3]] }
      qS <- qS + tmp[[6]] ; czS <- czS + tmp[[5]] ; cxS <- cxS + tmp[[7]]
      rm(fit, sm)
    }
    diagCkhiCChk[iMapThis,1] <- qS / length(iBlocksAdj)
    CkhiCX[iMapThis,] <- cxS / length(iBlocksAdj)
    CkhiCz_muhat[iMapThis,1] <- czS / length(iBlocksAdj)
  }
  gc(FALSE)
  return(list('diagCkk' = diagCkk , 'diagCkhiCChk' = diagCkhiCChk , 'CkhiCX' = CkhiCX , 'CkhiCz_muhat' = CkhiCz_muhat))
}

### ---- block construction (compositeLikIAK3D.R:803-1133): the distance / nearest-centroid passes on the device ----
getBlocksFromLocations <- function(xMap , compLikMats){
  vc <- do.call(rbind, lapply(compLikMats$listBlocks, function(b){ matrix(b$centroid, nrow = 1) }))
  .Call('iak3dR_nearest_centroid', as.matrix(xMap) + 0, vc + 0, FALSE)[[1]]
}
calcnPerCluster <- function(vcentres , xU){
  as.numeric(.Call('iak3dR_nearest_centroid', as.matrix(xU) + 0, as.matrix(vcentres) + 0, TRUE)[[2]])
}
### setVoronoiBlocks with its three slow steps replaced: greedy seeding (:853-884), Voronoi assignment (:887-888) and the
### deldir adjacency (:907-916); balanceNeighbours / balanceNeighbours2 / setVoronoiBlocksWrap of the reference call it unchanged.
setVoronoiBlocks <- function(x , nPerBlock = 50 , plotVor = F , vcentres = NULL){
  x <- as.matrix(x) + 0
  xU <- x[!duplicated(x),,drop=FALSE]
  if(is.null(vcentres)){ vcentres <- .Call('iak3dR_voronoi_seed', xU, as.integer(nPerBlock)) }
  nBlocks <- nrow(vcentres)
  iBlock <- .Call('iak3dR_nearest_centroid', xU, vcentres + 0, FALSE)[[1]]
  listBlocks <- list()
  for(i in 1:nBlocks){
    xUi <- xU[which(iBlock == i),,drop=FALSE]
    listBlocks[[i]] <- list(centroid = vcentres[i,,drop=FALSE], xU = xUi,
                            i = which(is.element(x[,1] , xUi[,1]) & is.element(x[,2] , xUi[,2])))
  }
  subsetPairsAdj <- .Call('iak3dR_dirichlet_adjacency', vcentres + 0)
  key <- function(m){ if(is.null(m) || nrow(m) == 0) character(0) else paste(m[,1], m[,2]) }
  allPairs <- if(nBlocks > 1) t(combn(nBlocks, 2)) else matrix(0L, 0, 2)
  subsetPairsNonadj <- allPairs[!(key(allPairs) %in% key(subsetPairsAdj)),,drop=FALSE]
  if(nrow(subsetPairsNonadj) == 0){ subsetPairsNonadj <- c() }
  return(list('listBlocks' = listBlocks , 'subsetPairsAdj' = subsetPairsAdj , 'subsetPairsNonadj' = subsetPairsNonadj))
}

### ---- makeXvX on the device for prediction grids (makeXvX.R:294-353, 'mlr' with lnTfmdData = FALSE) ----
### iak3dDesignMlr flattens modelX (+ allKnotsd, nDiscPts, XLims) into the specification of include/iak3d.h; Cubist designs are
### flattened the same way from cubistModel$listRules / listCoeffs_iv / comms4XCubist (iak3d_b200/design.py shows the layout).
iak3dDesignMlr <- function(namesX, covNames, covLevels = list(), allKnotsd = c(), nDiscPts = 10, XLims = NULL){
  if(length(allKnotsd) > 0){ namesX <- namesX[!(namesX %in% c('d', 'd2', 'd3'))] }
  p <- length(namesX) ; nint <- max(length(allKnotsd) - 2, 0)
  ipS <- which(substr(namesX, 1, 8) == 'dSpline.') ; nspl <- length(ipS)
  if(is.null(XLims)){ XL <- rbind(rep(-Inf, p), rep(Inf, p)) ; has <- 0L ; infB <- 1L
  }else{ XL <- XLims ; has <- 1L ; infB <- as.integer((min(XL[1,]) == -Inf) & (max(XL[2,]) == Inf)) }
  kind <- cov <- integer(p) ; level <- rep(NaN, p) ; l <- 2
  for(j in seq(p)){
    nm <- namesX[j] ; base <- NA
    if(nm == 'const'){ kind[j] <- 0L }else if(nm == 'd'){ kind[j] <- 1L }else if(nm == 'd2'){ kind[j] <- 2L }else if(nm == 'd3'){ kind[j] <- 3L
    }else if(substr(nm, 1, 8) == 'dSpline.'){ kind[j] <- 4L
    }else if(substr(nm, 1, 3) == 'd2.'){ kind[j] <- 2L ; base <- substr(nm, 4, nchar(nm))
    }else if(substr(nm, 1, 3) == 'd3.'){ kind[j] <- 3L ; base <- substr(nm, 4, nchar(nm))
    }else if(substr(nm, 1, 2) == 'd.'){ kind[j] <- 1L ; base <- substr(nm, 3, nchar(nm))
    }else{ kind[j] <- 0L ; base <- nm }
    if(kind[j] == 4L){ cov[j] <- match(j, ipS) - 1L
    }else if(is.na(base)){ cov[j] <- -1L
    }else{
      cov[j] <- match(base, covNames) - 1L
      if(!is.null(covLevels[[base]])){ if(j == 1 || namesX[j] != namesX[j-1]){ l <- 2 } ; level[j] <- l ; l <- l + 1 }
    }
  }
  ord <- c(which(namesX == 'd'), which(substr(namesX, 1, 2) == 'd.'), which(namesX == 'd2'), which(substr(namesX, 1, 3) == 'd2.'),
           which(namesX == 'd3'), which(substr(namesX, 1, 3) == 'd3.'))
  cpos <- rep(-1L, p) ; cpos[ord] <- seq_along(ord) - 1L
  ispec <- c(0L, p, length(covNames), nspl, 0L, nint, as.integer(nDiscPts), has, infB, as.integer(rbind(kind, cov, cpos)))
  dspec <- c(if(length(allKnotsd) > 0) allKnotsd[c(1, length(allKnotsd))] else c(0, 1), if(nint > 0) allKnotsd[2:(nint + 1)] else NULL,
             XL[1,], XL[2,], level)
  list(handle = .Call('iak3dR_design', as.integer(ispec), as.numeric(dspec)), namesX = namesX, covNames = covNames)
}
### predictIAK3D's double loop over depths and batches (predictIAK3D.R:144-255) in one call; returns list(zMap, vMap), ndIMap x nxMap
iak3dPredictGrid <- function(lmmFit, design, xMap, covsMap, dIMap, nxPerBatch = 2000){
  covM <- if(length(design$covNames) > 0) as.matrix(covsMap[, design$covNames, drop = FALSE]) + 0 else NULL
  tmp <- .Call('iak3dR_predict_grid', lmmFit$iak3dFit, design$handle, as.matrix(xMap) + 0, covM, round(as.matrix(dIMap), digits = 2), nxPerBatch)
  list(zMap = t(tmp[[1]]), vMap = t(tmp[[2]]))
}

### kriging of map supports with site-specific random effects (predictIAK3D.R:166-195): site labels of the map are matched
### against the levels the fit's setupMats recorded; unseen sites get new codes (no shared structure with any datum)
iak3dPredictBatchSsre <- function(lmmFit, xMapThis, dIMapThis, XMapThis, siteIDMapThis){
  lev <- lmmFit$setupMats$siteLevels ; lab <- as.character(siteIDMapThis)
  code <- match(lab, lev) ; new <- is.na(code) ; code[new] <- length(lev) + match(lab[new], unique(lab[new]))
  tmp <- .Call('iak3dR_predict_ssre', lmmFit$iak3dFit, .iak3d_dbl(xMapThis), .iak3d_dbl(dIMapThis), .iak3d_dbl(XMapThis),
               as.integer(code - 1L), .iak3d_dbl(as.matrix(XMapThis)[, lmmFit$setupMats$colnames4ssre, drop = FALSE]))
  list(z = tmp[[1]], v = tmp[[2]])
}
