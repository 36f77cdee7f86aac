// binom_fit.cuh -- one Binomial (logistic) EBEN fit per thread block.  (stub: filled in next)
#pragma once
#include "common.cuh"
namespace pareben {
template <bool EPIS>
__device__ void binom_fit(const Problem &P, const FoldData &F, const Variant &v, Slab &s, double lambda,
                          double alpha_en, const FitTask &task, const FitOutputs &out, double *sV, const Scratch &sc)
{
    if (threadIdx.x == 0) {
        if (out.fold_err) out.fold_err[task.out_index] = nan("");
        if (out.status) out.status[task.out_index] = ST_NONFINITE;
    }
}
}  // namespace pareben
