### iak3d_cuda.R -- drop-in GPU back end for the hot path of iak3d.
### source() this file AFTER the reference's own files (egEdgeroi.R:53-71) and after dyn.load("iak3dR.so")
### (r/r_shim.c).  It re-defines, with identical arguments and return values, the functions the optimiser and the
### predictor look up by name at call time:
###     setupIAK3D (adds a device handle), lndetANDinvCb, setCIAK3D, setCIAK3D2, nllIAK3D, gradnllIAK3D[.mc]
### and keeps everything else (readParsIAK3D, fitIAK3D, optimIt, makeXvX, ...) as the reference defines it.
### Forked workers cannot share a CUDA context: the glue forces the serial paths and batches those loops instead.

.iak3d_ref <- list(setupIAK3D = setupIAK3D, lndetANDinvCb = lndetANDinvCb, nllIAK3D = nllIAK3D)
.iak3d_modelx <- c(nugget = 0, spherical = 1, matern = 2, wendland = 3)

.iak3d_pv <- function(parsBTfmd, modelx, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, quirk_scale = 1){
  tau <- function(v, t){ if(t == -1){ as.numeric(v)[1:2] }else{ c(0, 0) } }
  nz <- function(v){ if(is.null(v) || length(v) == 0 || is.na(v)) NaN else as.numeric(v) }
  c(.iak3d_modelx[[modelx]], cmeOpt, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, 1,
    nz(parsBTfmd$ax), nz(parsBTfmd$nux), nz(parsBTfmd$ad), nz(parsBTfmd$nud),
    parsBTfmd$cx0, parsBTfmd$cx1, parsBTfmd$cd1, parsBTfmd$cxd0, parsBTfmd$cxd1, parsBTfmd$cme,
    tau(parsBTfmd$sdfdPars_cd1, sdfdType_cd1), tau(parsBTfmd$sdfdPars_cxd0, sdfdType_cxd0),
    tau(parsBTfmd$sdfdPars_cxd1, sdfdType_cxd1), quirk_scale)
}

### setupIAK3D: the id vectors live on the device; spline sd functions (sdfdType > 0) fall back to the reference.
setupIAK3D <- function(xData, dIData, ...){
  args <- list(...)
  types <- unlist(args[c('sdfdType_cd1', 'sdfdType_cxd0', 'sdfdType_cxd1')])
  if(length(types) > 0 && max(types) > 0){ return(.iak3d_ref$setupIAK3D(xData = xData, dIData = dIData, ...)) }
  xData <- as.matrix(xData) ; dIData <- as.matrix(dIData)
  list(iak3d = .Call('iak3dR_dataset', xData, dIData, NULL), xData = xData, dIData = dIData, n = nrow(xData),
       Kx = TRUE) # 'Kx' non-NULL: the reference's is.null(setupMats$Kx) guards (fitIAK3D.R:965) stay quiet
}

lndetANDinvCb <- function(C, b = NULL){
  C <- as.matrix(C)
  if(nrow(C) < 256){ return(.iak3d_ref$lndetANDinvCb(C, b)) } # the p x p systems stay on the host
  tmp <- .Call('iak3dR_lndet_invCb', C, if(is.null(b)) NULL else as.matrix(b))
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
  .Call('iak3dR_cov_cross', setupMats$set2$iak3d,
        .iak3d_pv(parsBTfmd, modelx, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt),
        as.matrix(setupMats$xData), as.matrix(setupMats$dIData))
}
setupIAK3D2 <- function(xData, dIData, xData2, dIData2, ...){
  list(xData = as.matrix(xData), dIData = as.matrix(dIData), set2 = setupIAK3D(xData2, dIData2), Kx = TRUE)
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
  if(lnTfmdData || is.null(setupMats$iak3d)){  # lognormal support / spline sd functions: reference path
    return(.iak3d_ref$nllIAK3D(pars, zData, XData, vXU, iU, modelx, nud, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, prodSum,
                               setupMats, parBnds, useReml, lnTfmdData, rtnAll, forCompLik, attachBigMats, nCores))
  }
  XData <- as.matrix(XData) ; n <- length(zData) ; p <- ncol(XData)
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
  XData <- as.matrix(XData) ; n <- length(zData) ; p <- ncol(XData) ; delta <- 1E-6
  .Call('iak3dR_set_W', setupMats$iak3d, cbind(XData, zData))
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
