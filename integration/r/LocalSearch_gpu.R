# Drop-in body for R/LocalSearch.R: same signature, same returned list (CrossValidation, alpha.optimal,
# lambda.optimal, fullCV) and the same visiting order.  The reference runs up to 400 sequential foreach barriers of
# nFolds fits each (R/LocalSearch.R:56-122); here the whole alpha x lambda x fold table is ONE .Call and the
# early-stopping walk is replayed over the returned means, so the rows the reference would never have visited
# are simply not read.  lambda_Max comes from the device as well (the K^2 R loop of :30-39).
LocalSearch <- function(BASIS, Target, nFolds, Epis = "no", foldId = 0, prior = "gaussian", nDevices = 0L){
  if(prior != "gaussian"){
    print("For the binomial prior, please use the global search.")   # R/LocalSearch.R:126-128
    return(invisible(NULL))
  }
  cat("EB-Elastic Net Linear Model, Epis: ", Epis, "", nFolds, "fold cross-validation\n")
  N <- nrow(BASIS)
  if(length(foldId) != N){                                            # R/LocalSearch.R:13-20 (unseeded, as there)
    if(N %% nFolds != 0){
      foldId <- sample(c(rep(1:nFolds, floor(N/nFolds)), 1:(N %% nFolds)), N)
    }else{
      foldId <- sample(rep(1:nFolds, floor(N/nFolds)), N)
    }
  }
  storage.mode(BASIS) <- "double"
  lambda_Max <- .Call("pareben_lambda_max_call", BASIS, as.double(Target), as.integer(Epis == "yes"), 0L,
                      PACKAGE = "parEBEN") * 10
  lambda_Min <- log(0.001 * lambda_Max)
  step <- (log(lambda_Max) - lambda_Min)/19
  Lambda <- exp(seq(from = log(lambda_Max), to = lambda_Min, by = -step))
  N_step <- length(Lambda)
  Alpha <- seq(from = 1, to = 0.05, by = -0.05)
  nAlpha <- length(Alpha)

  # every (alpha, lambda) x fold fit in one launch; column g = grid point g, alpha-major then lambda
  ga <- rep(Alpha, each = N_step); gl <- rep(Lambda, times = nAlpha)
  res <- .Call("pareben_cv_grid_call", BASIS, as.double(Target), as.integer(foldId), as.integer(nFolds),
               as.double(ga), as.double(gl), as.integer(Epis == "yes"), 0L, as.integer(nDevices), 0L, PACKAGE = "parEBEN")
  if(any(res$status != 0)) warning(sum(res$status != 0), " fits finished with a non-zero status")
  sse  <- colMeans(res$fold_err)
  se   <- apply(res$fold_err, 2, sd)/sqrt(nFolds)

  MSEcv <- mat.or.vec((N_step*nAlpha), 4)
  MSEeachAlpha <- mat.or.vec(nAlpha, 4)
  step <- 1
  for(i_alpha in 1:nAlpha){
    alpha <- Alpha[i_alpha]
    SSE1Alpha <- matrix(1e10, N_step, 2)
    for(i_s in 1:N_step){
      g <- (i_alpha - 1)*N_step + i_s
      min_index <- which.min(SSE1Alpha[1:(i_s - 1), 1])               # i_s = 1: rows 1:0 = (1, 0), as in the reference
      previousL <- SSE1Alpha[min_index, 1] + SSE1Alpha[min_index, 2]
      SSE1Alpha[i_s, ] <- c(sse[g], se[g])
      MSEcv[step, ] <- c(alpha, Lambda[i_s], sse[g], se[g])
      currentL <- MSEcv[step, 3]
      step <- step + 1
      if((currentL - previousL) > 0){break}
    }
    index <- which.min(SSE1Alpha[, 1])
    MSEeachAlpha[i_alpha, ] <- c(alpha, Lambda[index], SSE1Alpha[index, ])
  }
  index <- which.min(MSEeachAlpha[, 3])
  result <- list(MSEeachAlpha, MSEeachAlpha[index, 1], MSEeachAlpha[index, 2], MSEcv)
  names(result) <- c("CrossValidation", "alpha.optimal", "lambda.optimal", "fullCV")
  return(result)
}
