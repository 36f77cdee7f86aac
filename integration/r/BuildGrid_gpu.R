# Drop-in bodies for R/BuildGrid.R.  GetLambdaMax's loops -- K main effects and, for Epis = "yes", the K(K-1)/2 pairs
# scanned by a double for-loop in R (R/BuildGrid.R:21-30: 2e8 iterations at K = 20,000) -- become one device call;
# BuildGrid itself (R/BuildGrid.R:34-52) is unchanged.
GetLambdaMax <- function(BASIS, Target, Epis = "no"){
  storage.mode(BASIS) <- "double"
  .Call("pareben_lambda_max_call", BASIS, as.double(Target), as.integer(Epis == "yes"), 0L, PACKAGE = "parEBEN")
}

# Final model after cross-validation (README.md:89-94).  Either keep calling EBEN::EBelasticNet.Gaussian / .Binomial with
# EBEN's `useDynLib` pointed at libpareben.so -- it exports elasticNetLinearNeMainEff, elasticNetLinearNeEpisEff,
# ElasticNetBinaryNEmainEff and ElasticNetBinaryNEfull with the original argument lists -- or go through the .Call shim
# and apply EBEN's own post-processing (EBEN_orig/R/EBelasticNet.Gaussian.R:56-98) to the returned `.C` outputs:
EBelasticNet.Gaussian.gpu <- function(BASIS, Target, lambda, alpha, Epis = "no", verbose = 0){
  storage.mode(BASIS) <- "double"
  out <- .Call("pareben_fit_call", BASIS, as.double(Target), as.double(lambda), as.double(alpha),
               as.integer(Epis == "yes"), 0L, 0L, PACKAGE = "parEBEN")
  list(Beta = out$Beta, WaldScore = out$WaldScore, Intercept = out$Intercept, residual = out$extra)   # = the `output` list of :39-51
}
