// stream.cuh -- streaming-Kc mode: EBEN fits whose candidate set is too large for a per-fit cache
// (BASELINE config 5: K = 20,000 loci -> Kc = 200,010,000 main + pair candidates).
//
// The cached kernel (fit_kernel.cuh) keeps, per fit, the reference's BASIS_PHI cache (one row of Kc doubles per
// active basis, NeFull2.c:350-365) plus six Kc-length statistic arrays; at Kc = 2e8 that is 1.6 GB per active
// basis per fit and neither the reference nor that kernel can run.  Here a fit keeps ONLY its active set
// (PHI, SIGMA, mu, alpha, PHI'PHI) and the per-candidate statistics are recomputed from it whenever the solver
// asks for them:
//
//     Q_c = beta_s * phi_c' e'            e' = t - PHI mu - d          (one N-vector per fit)
//     S_c = beta_s - beta_s^2 g_c' SIGMA g_c,   g_c = PHI' phi_c
//
// which are the quantities the reference maintains incrementally (FullStat NeFull2.c:1063-1196 followed by the
// rank-1 corrections of the add / delete / re-estimate actions :1407-1640).  Two details make the recomputed
// values equal to the reference's to rounding, not just approximately:
//   * d accumulates mu'_j phi_j for every basis deleted since the last FullStat: the reference's delete
//     truncates the weight to an int (`int Mujj`, MainEff.c:1746) and drops the remainder, so its Q array is
//     consistent with a residual that still carries that remainder;
//   * beta_s is the noise precision of the last FullStat: a noise update that moves log(beta) by less than 1e-6
//     changes beta but not the statistic arrays (:716-727).
// (scripts/stream_equivalence.py demonstrates both on the CPU oracle.)
//
// An out-of-model candidate can only ever be ADDed, and only when a = S - Q^2 + 2 l1 + l2 < 0 (fEBDeltaML,
// NeFull2.c:1227-1404) -- impossible unless Q^2 > 2 l1 + l2 because S > 0.  So one pass needs Q for every
// candidate but S only for the few that pass this exact screen.  All fits that share a fold advance in
// lock-step rounds:
//
//   stream_advance_kernel   one block per fit: consume the previous scan (in-model statistics, decision,
//                           block of actions, noise update, outer-loop bookkeeping) up to the next point where
//                           the reference calls fEBDeltaML, then publish e' as one column of the fold's
//                           right-hand-side matrix E;
//   stream_scan_kernel      ONE dense FP64 contraction  X_fold' * E  for all waiting fits of every fold on the
//                           FP64 tensor cores (DMMA m8n8k4): candidate columns -- pair products x_i * x_j are
//                           generated in registers from the int8 loci -- against E tiles that a producer warp
//                           streams into shared memory with bulk-tensor copies (cp.async.bulk + mbarrier
//                           ring).  The epilogue screens, and for survivors computes S, the closed-form
//                           delta-ML and appends to the fit's ADD list; nothing of length Kc is ever stored.
#pragma once
#include "common.cuh"

namespace pareben {

enum : int { ST_LIST = 16 };                 // streaming mode: a fit's ADD list overflowed (mirrored in include/pareben.h)
enum : int { SP_START = 0, SP_WAIT = 1, SP_DONE = 2 };

constexpr int SK = 64;                       // rows per shared-memory stage of E
constexpr int SN = 64;                       // right-hand sides (fits) per tile
constexpr int SLD = SK + 4;                  // leading dimension of a staged column: == 4 (mod 16) doubles -> conflict-free B fragments
constexpr int STAGE_D = SN * SLD;            // doubles per stage (34,816 bytes: one bulk copy)
constexpr int SCAN_STAGES = 4;
constexpr int SCAN_CT = 128;                 // candidates per work item: 8 consumer warps x 16
constexpr int SCAN_WARPS = 8;                // consumer warps (one more warp is the copy producer)
constexpr int SCAN_THREADS = 32 * (SCAN_WARPS + 1);
constexpr int STREAM_MAX_FOLDS = 256;
constexpr int ADV_THREADS = 256;

struct StreamFit {
    // task
    int fold, out_index;
    double lambda, alpha_en;
    // control state of the solver between two rounds (the loop-carried scalars of gauss_fit())
    int phase;
    int iter, i_iter, selected, ini_removed, n_update, jj, it_max, initial, M, status;
    double beta, beta_s, b, vk, vk0, err, residvar, var_y, flops;
    // scan interface (meaningful while phase == SP_WAIT)
    int slot;                               // right-hand-side slot within the fold this round
    double l1, l2, ml_delta, n_add;
    unsigned long long runmax;              // bit pattern of the largest ADD delta-ML seen so far this round (0 = none)
    int n_list, any_add;
    // per-fit arrays
    double *sigma, *sigma_new, *H, *ptp, *phi;
    double *mu, *alpha, *gamma, *tmp, *u, *colk, *ascale, *s_in, *q_in, *dml_in, *aroot_in;
    double *t, *e, *d, *phinew;
    int *used, *act_in, *blk_c, *blk_src;
    int *list_c; double *list_dml, *list_aroot;
};

// Per-fold view of the lock-step round.
struct StreamFold {
    double *E;                  // [rhs tile][stage][SN][SLD] staged right-hand sides of this round
    double *thr;                // [slot] screen: z_raw^2 > thr * ssq_c
    int *slot_fit;              // [slot] fit index
    int max_slots;
};

struct StreamShared {
    StreamFold *folds;          // [n_folds + 1]
    int *n_slots;               // [n_folds + 1] slots handed out this round (advance kernel: atomicAdd; host: reset)
    int list_cap;
    double *warp_scratch;       // [scan blocks][SCAN_WARPS][cap]
    double *flops;
};

__host__ __device__ inline size_t e_tile_doubles(int ldt) { return (size_t)(ldt / SK) * STAGE_D; }

}  // namespace pareben
