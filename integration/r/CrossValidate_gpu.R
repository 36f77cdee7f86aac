# Drop-in body for R/CrossValidate.R (search = "global"): the two foreach/%dopar% loops
# (R/CrossValidate.R:66-70 and :88-92) become one .Call.  Signature, return list and column names
# are unchanged.  Folds: TestModel ignores the caller's foldId and recomputes them with
# set.seed(1) inside every task (R/TestModel.R:9), so the same is done here, once.
CrossValidate <- function(BASIS, Target, nFolds, foldId = 0, Epis = "no", prior = "gaussian", search = "global", nDevices = 0L){
  if(search == "global"){
    ParameterGrid <- BuildGrid(BASIS, Target, nFolds, Epis)
    folds <- AssignToFolds(BASIS, nFolds)
    storage.mode(BASIS) <- "double"
    res <- .Call("pareben_cv_grid_call", BASIS, as.double(Target), as.integer(folds), as.integer(nFolds),
                 as.double(ParameterGrid$alpha), as.double(ParameterGrid$lambda),
                 as.integer(Epis == "yes"), as.integer(prior != "gaussian"), as.integer(nDevices), 0L, PACKAGE = "parEBEN")   # nDevices = 0: every GPU of the box
    if(any(res$status != 0)) warning(sum(res$status != 0), " fits finished with a non-zero status")
    detail <- data.frame(foldId = rep(1:nFolds, nrow(ParameterGrid)),
                         alpha  = rep(ParameterGrid$alpha,  each = nFolds),
                         lambda = rep(ParameterGrid$lambda, each = nFolds))
    if(prior == "gaussian"){
      detail$MSE <- as.vector(res$fold_err)
      Error <- detail %>% group_by(alpha, lambda) %>%
        summarise(SE = sd(MSE)/sqrt(max(foldId)), MSE = mean(MSE))
      index <- which.min(Error$MSE)
    }else{
      detail$logL <- as.vector(res$fold_err)
      Error <- detail %>% group_by(alpha, lambda) %>%
        summarise(SE = sd(logL)/sqrt(max(foldId)), Likelihood = -mean(logL))
      index <- which.min(Error$Likelihood)   # the reference indexes Error$MSE here and returns empty (R/CrossValidate.R:99)
    }
    list(Results.Detail = detail, Results.Summary = Error,
         lambda.optimal = Error[index,]$lambda, alpha.optimal = Error[index,]$alpha)
  }else{
    LocalSearch(BASIS, Target, nFolds, Epis, foldId, prior, nDevices)   # integration/r/LocalSearch_gpu.R
  }
}
