/* TEST INFRASTRUCTURE -- CPU restatement (oracle "port") of EBEN's Gaussian empirical-Bayes
 * elastic-net fit, main-effect and Epis variants.  Plain C, no BLAS, single thread.
 *
 * Restates (file:line under /root/reference/EBEN_orig/src/):
 *   elasticNetLinearNeMainEff.c  entry 55-242, inner solver 248-809, init 976-1108,
 *                                CacheBP 1144-1201, FullStat 1209-1341, DeltaML 1372-1582,
 *                                ActionAdd 1585-1723, ActionDel 1725-1822, FinalUpdate 1841-1921
 *   elasticNetLinearNeFull2.c    same skeleton with K(K+1)/2 candidates (entry 57-261) and the
 *                                constants that differ (SURVEY.md Appendix A "variant constant table").
 * Parity is pinned by tests/test_oracle.py (test_port_*) against oracle/_ref/libeben_ref.so (the
 * reference C compiled unmodified here) and against tests/golden/ vectors generated from it;
 * summation order differs from OpenBLAS inside ddot/dgemv/dgemm/dpotrf, so agreement is to
 * rounding (rel <= 1e-9 on fold errors, identical supports), not bit-for-bit.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may link this file.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { ACT_REEST = 0, ACT_ADD = 1, ACT_DEL = -1, ACT_TERM = 10, ACT_NONE = -10 };

/* Test switch (tests/test_oracle.py::test_streaming_identity, scripts/stream_equivalence.py): when set, the statistic
 * arrays S_in / Q_in are NOT corrected incrementally by the actions but recomputed after every action from the
 * active set alone,
 *     S_c = beta_s - beta_s^2 g_c' SIGMA g_c,     Q_c = beta_s (x_c' t / s_c - g_c' mu) - (ghost term),
 * beta_s = the noise precision of the last FullStat, ghost term = beta * G[j][c] * mu'_j summed over the bases j deleted
 * since then (the remainder the reference's `int Mujj` truncation leaves behind, MainEff.c:1746).  This is the identity
 * the streaming kernels (pareben_b200/csrc/stream.cuh) are built on; the switch exists to demonstrate it against the
 * unmodified path on the CPU. */
static int g_fresh_statistics = 0;
void oracle_set_fresh_statistics(int on) { g_fresh_statistics = on; }

typedef struct {
    int epis;              /* 0 main, 1 pairwise */
    double n_add;          /* block threshold factor      MainEff.c:275 / NeFull2.c:293 */
    double ml_delta;       /* minimum worthwhile dML      MainEff.c:277 / NeFull2.c:295 */
    double reest_tol;      /* |dlog alpha| stop           MainEff.c:543 / NeFull2.c:558 */
    double init_alpha_max; /*                             MainEff.c:995 / NeFull2.c:849 */
} Variant;

typedef struct {
    int N, K, Kc, cap;
    const double *X;       /* N x K column-major */
    const int *loc1, *loc2; /* 0-based loci of candidate c */
    double *scale;         /* Kc */
    /* active-set state, persists across outer iterations */
    int M;
    int *used;             /* cap, 1-based candidate ids in insertion order */
    int *unused;           /* Kc */
    int n_unused;
    double *alpha, *mu, *gamma; /* cap */
    double *sigma, *sigma_new, *H; /* cap*cap, column-major with leading dim M */
    double *phi;           /* N x cap, normalised active columns */
    double **G;            /* cap rows of Kc: G[j][c] = phi_j' x_c / s_c  (BASIS_PHI) */
    double *xt;            /* Kc: x_c' t / s_c (BASIS_Targets) */
    double *S_in, *Q_in, *S_out, *Q_out, *dml, *aroot;
    int *action, *block;
    double beta;
    int overflow;
    double beta_stat;      /* fresh-statistics test switch: beta of the last FullStat */
    double *qcorr;         /* Kc: ghost term of the deleted bases since the last FullStat */
} State;

static void column(const State *s, int c, double *out)
{   /* raw (unscaled) candidate column: main BASIS[,c]; pair BASIS[,i]*BASIS[,j] (NeFull2.c:277-290) */
    const double *a = s->X + (size_t)s->loc1[c] * s->N;
    if (s->loc1[c] == s->loc2[c]) { memcpy(out, a, sizeof(double) * s->N); return; }
    const double *b = s->X + (size_t)s->loc2[c] * s->N;
    for (int h = 0; h < s->N; h++) out[h] = a[h] * b[h];
}

static double var_targets(const double *t, int n)
{   /* MainEff.c:1826-1838 */
    double m = 0, v = 0;
    for (int i = 0; i < n; i++) m += t[i];
    m /= n;
    for (int i = 0; i < n; i++) v += (t[i] - m) * (t[i] - m);
    return v / (n - 1);
}

/* Symmetric positive-definite inverse through an upper Cholesky factor, the role dpotrf+dpotri
 * play at MainEff.c:1346-1369.  a is n x n, leading dimension n.  Returns non-zero when the
 * matrix is not positive definite (the reference then leaves garbage and carries on). */
static int spd_inverse(double *a, int n)
{
    for (int j = 0; j < n; j++) {              /* a = U'U, U stored in the upper triangle (col-major: a[j*n+i], i<=j) */
        double d = a[j * n + j];
        for (int k = 0; k < j; k++) d -= a[j * n + k] * a[j * n + k];
        if (!(d > 0)) return j + 1;
        d = sqrt(d);
        a[j * n + j] = d;
        for (int i = j + 1; i < n; i++) {
            double v = a[i * n + j];
            for (int k = 0; k < j; k++) v -= a[i * n + k] * a[j * n + k];
            a[i * n + j] = v / d;
        }
    }
    /* U^-1 in place (upper) */
    for (int j = 0; j < n; j++) {
        a[j * n + j] = 1.0 / a[j * n + j];
        for (int i = 0; i < j; i++) {
            double v = 0;
            for (int k = i; k < j; k++) v += a[k * n + i] * a[j * n + k];
            a[j * n + i] = -v * a[j * n + j];
        }
    }
    /* A^-1 = U^-1 U^-T, upper triangle, then mirror */
    for (int i = 0; i < n; i++)
        for (int j = i; j < n; j++) {
            double v = 0;
            for (int k = j; k < n; k++) v += a[k * n + i] * a[k * n + j];
            a[j * n + i] = v;
        }
    for (int i = 1; i < n; i++)
        for (int j = 0; j < i; j++) a[j * n + i] = a[i * n + j];
    return 0;
}

static void rebuild_unused(State *s)
{   /* MainEff.c:1092-1107 */
    char *isused = calloc(s->Kc + 1, 1);
    for (int j = 0; j < s->M; j++) isused[s->used[j]] = 1;
    int kk = 0;
    for (int c = 0; c < s->Kc; c++) if (!isused[c + 1]) s->unused[kk++] = c + 1;
    s->n_unused = kk;
    free(isused);
}

static void initialise(State *s, const Variant *v, const double *t)
{   /* MainEff.c:1003-1090 / NeFull2.c:857-945.  The "least correlated" scan compares
     * fabs(proj) < fabs(0) and therefore never fires: the first basis is always candidate 1. */
    int N = s->N;
    s->M = 1;
    s->used[0] = 1;
    double sc = s->scale[0];
    for (int h = 0; h < N; h++) s->phi[h] = s->X[h] * (1 / sc);
    double var = var_targets(t, N);
    if (!v->epis) s->beta = 1 / (var * 0.01 + 1e-10);                     /* MainEff.c:1068 */
    else { double sd = sqrt(var); if (sd < 1e-6) sd = 1e-6; s->beta = 1 / pow(sd * 0.1, 2); } /* NeFull2.c:917-921 */
    double p = 0, q = 0;
    for (int h = 0; h < N; h++) { p += s->phi[h] * s->phi[h]; q += s->phi[h] * t[h]; }
    p *= s->beta; q *= s->beta;
    double a = p * p / (q * q - p);
    if (a < 0) a = v->init_alpha_max;
    if (a > v->init_alpha_max) a = v->init_alpha_max;
    s->alpha[0] = a;
}

static void cache_bp(State *s, const double *t)
{   /* CacheBP*: G = PHI' X / s, xt = X't / s  (MainEff.c:1157-1178, NeFull2.c:1010-1050) */
    int N = s->N;
    double *col = malloc(sizeof(double) * N);
    for (int c = 0; c < s->Kc; c++) {
        column(s, c, col);
        for (int l = 0; l < s->M; l++) {
            const double *ph = s->phi + (size_t)l * N;
            double z = 0;
            for (int h = 0; h < N; h++) z += ph[h] * col[h];
            s->G[l][c] = z / s->scale[c];
        }
        double z = 0;
        for (int h = 0; h < N; h++) z += col[h] * t[h];
        s->xt[c] = z / s->scale[c];
    }
    free(col);
}

static void refresh_out(State *s)
{   /* S_out/Q_out: copy, then in-model correction (MainEff.c:664-671, 1320-1338) */
    memcpy(s->S_out, s->S_in, sizeof(double) * s->Kc);
    memcpy(s->Q_out, s->Q_in, sizeof(double) * s->Kc);
    for (int i = 0; i < s->M; i++) {
        int c = s->used[i] - 1;
        s->S_out[c] = s->alpha[i] * s->S_in[c] / (s->alpha[i] - s->S_in[c]);
        s->Q_out[c] = s->alpha[i] * s->Q_in[c] / (s->alpha[i] - s->S_in[c]);
    }
}

static void posterior_mean(State *s, const double *t)
{   /* Mu = beta * SIGMA * PHI' t  (MainEff.c:1256-1280, 1894-1916) */
    int N = s->N, M = s->M;
    double *pt = malloc(sizeof(double) * M);
    for (int i = 0; i < M; i++) {
        double z = 0;
        for (int h = 0; h < N; h++) z += s->phi[(size_t)i * N + h] * t[h];
        pt[i] = z;
    }
    for (int i = 0; i < M; i++) {
        double z = 0;
        for (int j = 0; j < M; j++) z += s->sigma[j * M + i] * pt[j];
        s->mu[i] = z * s->beta;
    }
    free(pt);
}

static void full_stat(State *s, const double *t, int first)
{   /* fEBLinearFullStat* (MainEff.c:1209-1341) */
    int N = s->N, M = s->M;
    if (first) {   /* :1236-1246 */
        double h = 0;
        for (int k = 0; k < N; k++) h += s->phi[k] * s->phi[k];
        s->H[0] = h * s->beta + s->alpha[0];
        s->sigma[0] = 1 / s->H[0];
    }
    s->beta_stat = s->beta;
    if (s->qcorr) memset(s->qcorr, 0, sizeof(double) * s->Kc);
    posterior_mean(s, t);
    for (int i = 1; i < M; i++) s->gamma[i] = 1 - s->sigma[i * M + i] * s->alpha[i];   /* gamma[0] skipped, :1283 */
    double *bp = malloc(sizeof(double) * M);
    for (int c = 0; c < s->Kc; c++) {
        for (int j = 0; j < M; j++) {
            double z = 0;
            for (int p = 0; p < M; p++) z += s->G[p][c] * s->sigma[j * M + p];
            bp[j] = z;
        }
        double quad = 0, gm = 0;
        for (int j = 0; j < M; j++) quad += bp[j] * s->G[j][c];
        for (int p = 0; p < M; p++) gm += s->G[p][c] * s->mu[p];
        s->S_in[c] = s->beta - s->beta * quad * s->beta;
        s->Q_in[c] = s->beta * (s->xt[c] - gm);
    }
    free(bp);
    refresh_out(s);
}

static void final_update(State *s, const double *t)
{   /* FinalUpdate*: H = beta PHI'PHI + diag(alpha); SIGMA = H^-1; Mu (MainEff.c:1841-1921) */
    int N = s->N, M = s->M;
    for (int i = 0; i < M; i++)
        for (int j = 0; j < M; j++) {
            double z = 0;
            const double *a = s->phi + (size_t)i * N, *b = s->phi + (size_t)j * N;
            for (int h = 0; h < N; h++) z += a[h] * b[h];
            s->H[j * M + i] = z * s->beta;
        }
    for (int i = 0; i < M; i++) s->H[i * M + i] += s->alpha[i];
    memcpy(s->sigma, s->H, sizeof(double) * M * M);
    spd_inverse(s->sigma, M);
    posterior_mean(s, t);
}

typedef struct { double max; int nu; int any_delete; } Decision;

static Decision delta_ml(State *s, const Variant *v, double lambda, double alpha_en, double residual,
                         double var_y, int iter, int i_iter)
{   /* fEBDeltaML* (MainEff.c:1372-1582; NeFull2.c:1227-1404) */
    const double l1 = lambda * alpha_en, l2 = lambda * (1 - alpha_en);
    int M = s->M, Kc = s->Kc, any_add = 0, any_del = 0;
    int prio_add = 0, prio_del = 0;
    if (M < 10) { prio_add = 1; prio_del = 0; }
    if (M > 100 || (!v->epis && M >= s->N) || residual <= var_y * 0.1) { prio_add = 0; prio_del = 1; }
    for (int c = 0; c < Kc; c++) s->action[c] = ACT_NONE;
    double best = 0; int arg = 0;
    for (int pass = 0; pass < 2; pass++) {
        int n = pass == 0 ? M : s->n_unused;
        for (int i = 0; i < n; i++) {
            int c = (pass == 0 ? s->used[i] : s->unused[i]) - 1;
            double so = s->S_out[c], qo = s->Q_out[c];
            s->dml[c] = 0;
            double a = so - qo * qo + 2 * l1 + l2;
            double b = (so + l2) * (so + 4 * l1 + l2);
            double g = 2 * l1 * (so + l2) * (so + l2);
            double d = b * b - 4 * a * g;
            if (a < 0 && d > 0) {
                double r = (-b - sqrt(d)) / (2 * a);
                double L = (log(r / (r + so + l2)) + pow(qo, 2) / (r + so + l2)) * 0.5 - l1 / r;
                if (L > 0) {
                    s->aroot[c] = r + l2;
                    if (pass == 0) {
                        s->action[c] = ACT_REEST;
                        double o = s->alpha[i] - l2;
                        s->dml[c] = 0.5 * (log(r * (o + so + l2) / (o * (r + so + l2)))
                                           + qo * qo * (1 / (r + so + l2) - 1 / (o + so + l2)))
                                    - l1 * (1 / r - 1 / o);
                    } else {
                        s->action[c] = ACT_ADD;
                        s->dml[c] = L;
                        if (!v->epis) any_add = 1;      /* only the main-effect file sets it (:1484) */
                    }
                }
            } else if (pass == 0 && M > 1) {
                any_del = 1;
                s->action[c] = ACT_DEL;
                double o = s->alpha[i] - l2;
                double L = (log(o / (o + so + l2)) + pow(qo, 2) / (o + so + l2)) * 0.5 - l1 / o;
                s->dml[c] = -L;
            }
            if (s->dml[c] > best) { best = s->dml[c]; arg = c; }
        }
    }
    int rescan = 0;
    if ((any_add && prio_add) || (any_del && prio_del)) {      /* :1527-1556 */
        for (int c = 0; c < Kc; c++) {
            if (s->action[c] == ACT_REEST) s->dml[c] = 0;
            else if (s->action[c] == ACT_DEL) { if (any_add && prio_add && !prio_del) s->dml[c] = 0; }
            else if (s->action[c] == ACT_ADD) { if (any_del && prio_del && !prio_add) s->dml[c] = 0; }
        }
        rescan = 1;
    }
    if (rescan) { best = 0; arg = 0; for (int c = 0; c < Kc; c++) if (s->dml[c] > best) { best = s->dml[c]; arg = c; } }
    if (!v->epis && ((!any_add && iter == 1 && i_iter < 10) || (!any_add && residual >= var_y * 0.95))) { /* :1557-1577 */
        for (int c = 0; c < Kc; c++) if (s->action[c] == ACT_DEL) s->dml[c] = 0;
        best = 0; arg = 0;
        for (int c = 0; c < Kc; c++) if (s->dml[c] > best) { best = s->dml[c]; arg = c; }
    }
    Decision out = { best, arg, any_del };
    return out;
}

static void action_reestimate(State *s, int jj, double new_alpha)
{   /* MainEff.c:553-596 */
    int M = s->M;
    double old = s->alpha[jj];
    s->alpha[jj] = new_alpha;
    double kappa = 1.0 / (s->sigma[jj * M + jj] + 1.0 / (new_alpha - old));
    double mujj = s->mu[jj];
    const double *sj = s->sigma + jj * M;
    for (int i = 0; i < M; i++) s->mu[i] += -mujj * kappa * sj[i];
    for (int i = 0; i < M; i++)
        for (int j = 0; j < M; j++) s->sigma_new[j * M + i] = s->sigma[j * M + i] - kappa * sj[i] * sj[j];
    for (int c = 0; c < s->Kc; c++) {
        double z = 0;
        for (int j = 0; j < M; j++) z += s->G[j][c] * sj[j];
        s->S_in[c] += pow(s->beta * z, 2) * kappa;
        s->Q_in[c] += s->beta * mujj * kappa * z;
    }
}

static void action_add(State *s, int nu, double new_alpha, const double *phi_new)
{   /* ActionAdd* (MainEff.c:1585-1723; NeFull2.c:1407-1547) */
    int N = s->N, M = s->M, M1 = M + 1;
    double *g_new = malloc(sizeof(double) * s->Kc);
    double *col = malloc(sizeof(double) * N);
    for (int c = 0; c < s->Kc; c++) {
        column(s, c, col);
        double z = 0;
        for (int h = 0; h < N; h++) z += col[h] * phi_new[h];
        g_new[c] = z / s->scale[c];
    }
    free(col);
    double *tmp = malloc(sizeof(double) * M), *u = malloc(sizeof(double) * M);
    for (int i = 0; i < M; i++) {
        double z = 0;
        for (int h = 0; h < N; h++) z += s->phi[(size_t)i * N + h] * phi_new[h];
        tmp[i] = z * s->beta;
    }
    for (int i = 0; i < M; i++) {
        double z = 0;
        for (int j = 0; j < M; j++) z += s->sigma[i * M + j] * tmp[j];
        u[i] = z;
    }
    s->alpha[M] = new_alpha;
    memcpy(s->phi + (size_t)M * N, phi_new, sizeof(double) * N);
    double s_ii = 1.0 / (new_alpha + s->S_in[nu]);
    double mu_i = s_ii * s->Q_in[nu];
    for (int i = 0; i < M; i++) s->mu[i] += -mu_i * u[i];
    s->mu[M] = mu_i;
    for (int i = 0; i < M; i++)
        for (int j = 0; j < M; j++) s->sigma_new[j * M1 + i] = s->sigma[j * M + i] + (s_ii * u[i]) * u[j];
    for (int i = 0; i < M; i++) { s->sigma_new[M * M1 + i] = -s_ii * u[i]; s->sigma_new[i * M1 + M] = -s_ii * u[i]; }
    s->sigma_new[M * M1 + M] = s_ii;
    for (int c = 0; c < s->Kc; c++) {
        double z = 0;
        for (int j = 0; j < M; j++) z += s->G[j][c] * u[j];
        double mci = s->beta * g_new[c] - s->beta * z;
        s->S_in[c] -= mci * mci * s_ii;
        s->Q_in[c] -= mu_i * mci;
    }
    s->G[M] = g_new;
    free(tmp); free(u);
}

static void action_delete(State *s, int jj)
{   /* ActionDel* (MainEff.c:1725-1822).  Note `int Mujj`: the weight is truncated toward zero. */
    int N = s->N, M = s->M, last = M - 1;
    s->alpha[jj] = s->alpha[last];
    memmove(s->phi + (size_t)jj * N, s->phi + (size_t)last * N, sizeof(double) * N);
    int mujj = (int)s->mu[jj];
    const double *sj = s->sigma + jj * M;
    double sjj = sj[jj];
    for (int i = 0; i < M; i++) s->mu[i] = s->mu[i] - mujj * sj[i] / sjj;
    if (s->qcorr) { const double left = s->mu[jj]; for (int c = 0; c < s->Kc; c++) s->qcorr[c] += s->beta * s->G[jj][c] * left; }
    s->mu[jj] = s->mu[last];
    double *tmp = malloc(sizeof(double) * M * M);
    for (int i = 0; i < M; i++)
        for (int j = 0; j < M; j++) tmp[j * M + i] = s->sigma[j * M + i] - sj[i] / sjj * sj[j];
    for (int i = 0; i < last; i++)
        for (int j = 0; j < last; j++) s->sigma_new[j * last + i] = tmp[j * M + i];
    if (jj != last) {   /* move the last row/column into slot jj (:1763-1776) */
        for (int i = 0; i < last; i++) s->sigma_new[jj * last + i] = tmp[last * M + i];
        tmp[jj * M + M - 1] = tmp[M * M - 1];
        for (int i = 0; i < last; i++) s->sigma_new[i * last + jj] = tmp[i * M + M - 1];
    }
    free(tmp);
    for (int c = 0; c < s->Kc; c++) {
        double z = 0;
        for (int j = 0; j < M; j++) z += s->G[j][c] * sj[j];
        s->S_in[c] += pow(s->beta * z, 2) / sjj;
        s->Q_in[c] += s->beta * z * mujj / sjj;
    }
    double *p = s->G[jj]; s->G[jj] = s->G[last]; s->G[last] = p;
}

static void fresh_statistics(State *s)
{   /* the streaming identity: S_in, Q_in from (SIGMA_new, mu, the cache, beta_stat, ghost term) alone */
    int M = s->M;
    const double *sg = s->sigma_new, bs = s->beta_stat;
    double *bp = malloc(sizeof(double) * M);
    for (int c = 0; c < s->Kc; c++) {
        for (int j = 0; j < M; j++) { double z = 0; for (int p = 0; p < M; p++) z += s->G[p][c] * sg[j * M + p]; bp[j] = z; }
        double quad = 0, gm = 0;
        for (int j = 0; j < M; j++) quad += bp[j] * s->G[j][c];
        for (int p = 0; p < M; p++) gm += s->G[p][c] * s->mu[p];
        s->S_in[c] = bs - bs * quad * bs;
        s->Q_in[c] = bs * (s->xt[c] - gm) - s->qcorr[c];
    }
    free(bp);
}

static void after_action(State *s)
{   /* MainEff.c:657-681 */
    if (g_fresh_statistics) fresh_statistics(s);
    refresh_out(s);
    memcpy(s->sigma, s->sigma_new, sizeof(double) * s->M * s->M);
    for (int i = 0; i < s->M; i++) s->gamma[i] = 1 - s->alpha[i] * s->sigma[i * s->M + i];
}

/* One call of LinearFastEmpBayes* (MainEff.c:248-809) minus the N x N C_inv, which the caller
 * only ever uses through 1'C^-1 1 and 1'C^-1 y. */
static void inner_solver(State *s, const Variant *v, const double *t, double lambda, double alpha_en,
                         int iter, double residual, double var_y)
{
    int N = s->N, Kc = s->Kc;
    int ini_removed = 1, initial;
    if (iter <= 1) { initialise(s, v, t); ini_removed = 0; }
    rebuild_unused(s);
    memset(s->gamma, 0, sizeof(double) * s->cap);   /* gamma is Calloc'ed afresh per call (:344); gamma[0] stays 0 until an action */
    initial = s->used[0];
    for (int i = 0; i < s->M; i++) if (!s->G[i]) s->G[i] = malloc(sizeof(double) * Kc);
    cache_bp(s, t);
    int i_iter = 0;
    full_stat(s, t, iter == 1);
    int selected = ACT_NONE, last = 0, n_update = 0, jj = -1;
    int it_max = iter == 1 ? 10 : 100;
    double *phi_new = malloc(sizeof(double) * N), *e = malloc(sizeof(double) * N);
    while (!last) {
        i_iter++;
        Decision d = delta_ml(s, v, lambda, alpha_en, residual, var_y, iter, i_iter);
        int nu = d.nu, worthwhile;
        if (selected == ACT_TERM && !ini_removed && s->M > 1) nu = -1;        /* :426-430 */
        if (nu == -1 && ini_removed) { worthwhile = 0; selected = ACT_TERM; }
        else if (nu == -1 && !ini_removed && s->M > 1) {                       /* :437-446 */
            worthwhile = 1; nu = initial - 1;
            s->action[nu] = ACT_DEL; n_update = 1; s->block[0] = nu; ini_removed = 1; selected = ACT_DEL;
        } else {
            worthwhile = 1;
            double cutoff = d.max * (s->action[nu] == ACT_ADD ? v->n_add : 1.0);
            if (cutoff < v->ml_delta) cutoff = v->ml_delta;
            n_update = 0;
            for (int c = 0; c < Kc; c++) if (s->dml[c] >= cutoff) s->block[n_update++] = c;
            if (s->action[nu] == ACT_DEL && n_update > 1) n_update = 1;
            if (n_update == 0) worthwhile = 0;
        }
        if (!worthwhile) selected = ACT_TERM;
        if (worthwhile) {
            for (int iu = 0; iu < n_update; iu++) {
                nu = s->block[iu];
                selected = s->action[nu];
                double new_alpha = s->aroot[nu];
                if (selected == ACT_REEST || selected == ACT_DEL)
                    for (int i = 0; i < s->M; i++) if (s->used[i] == nu + 1) { jj = i; break; }
                column(s, nu, phi_new);
                if (s->loc1[nu] == s->loc2[nu]) for (int h = 0; h < N; h++) phi_new[h] *= 1 / s->scale[nu];   /* dscal by 1/s (:516-520) */
                else for (int h = 0; h < N; h++) phi_new[h] = phi_new[h] / s->scale[nu];                       /* pairs: x/s (NeFull2.c:277-290) */
                if (selected == ACT_REEST && fabs(log(new_alpha) - log(s->alpha[jj])) <= v->reest_tol && !d.any_delete)
                    selected = ACT_TERM;
                int updated = 0;
                if (selected == ACT_REEST) { action_reestimate(s, jj, new_alpha); updated = 1; }
                else if (selected == ACT_ADD) {
                    if (s->M + 1 > s->cap) { s->overflow = 1; selected = ACT_TERM; }   /* reference overruns its heap here (:605-611) */
                    else {
                        action_add(s, nu, new_alpha, phi_new);
                        s->used[s->M] = nu + 1;
                        s->n_unused--;
                        for (int i = 0; i < s->n_unused; i++) if (s->unused[i] == nu + 1) s->unused[i] = s->unused[s->n_unused];
                        s->M++;
                        updated = 1;
                    }
                } else if (selected == ACT_DEL) {
                    action_delete(s, jj);
                    int lastj = s->M - 1;
                    free(s->G[lastj]); s->G[lastj] = NULL;
                    s->used[jj] = s->used[lastj];
                    s->unused[s->n_unused++] = nu + 1;
                    s->M--;
                    updated = 1;
                }
                if (updated) after_action(s);
            }
        }
        if (selected == ACT_TERM || i_iter <= 10 || i_iter % 5 == 0 || n_update >= 2) {   /* :685-729 */
            int M = s->M;
            double ee = 0;
            for (int h = 0; h < N; h++) {
                double pm = 0;
                for (int j = 0; j < M; j++) pm += s->phi[(size_t)j * N + h] * s->mu[j];
                e[h] = t[h] - pm;
                ee += e[h] * e[h];
            }
            double beta_old = s->beta, sg = 0;
            for (int i = 0; i < M; i++) sg += s->gamma[i];
            s->beta = (N - sg) / ee;
            double vt = var_targets(t, N);
            if (s->beta > 1e6 / vt) s->beta = 1e6 / vt;
            if (fabs(log(s->beta) - log(beta_old)) > 1e-6) {
                final_update(s, t);
                if (selected != ACT_TERM) full_stat(s, t, 0);
            }
        }
        if (selected == ACT_TERM && ini_removed) last = 1;
        if ((i_iter == it_max && s->M == 1) || i_iter > it_max) last = 1;
        if (i_iter == it_max) selected = ACT_TERM;
    }
    free(phi_new); free(e);
}

static void fit(const Variant *v, const double *X, const double *y, double lambda, double alpha_en,
                double *Beta, double *wald, double *intercept, int N, int K, double *residual)
{
    State s; memset(&s, 0, sizeof s);
    int Kc = v->epis ? (K + 1) * K / 2 : K;
    s.N = N; s.K = K; s.Kc = Kc; s.X = X;
    if (!v->epis) { s.cap = (int)(1e7 / Kc); if (s.cap > Kc) s.cap = Kc; }    /* MainEff.c:68-69 */
    else s.cap = N > K ? 2 * K : (N < 200 ? 4 * K : K);                        /* NeFull2.c:67-80 */
    int *loc1 = malloc(sizeof(int) * Kc), *loc2 = malloc(sizeof(int) * Kc);
    for (int i = 0; i < K; i++) loc1[i] = loc2[i] = i;
    if (v->epis) { int kk = K; for (int i = 0; i < K - 1; i++) for (int j = i + 1; j < K; j++) { loc1[kk] = i; loc2[kk] = j; kk++; } }
    s.loc1 = loc1; s.loc2 = loc2;
    s.scale = malloc(sizeof(double) * Kc);
    int ncol = v->epis ? 5 : 4;
    double *col = malloc(sizeof(double) * N);
    for (int c = 0; c < Kc; c++) {
        Beta[c] = loc1[c] + 1; Beta[Kc + c] = loc2[c] + 1;
        for (int k = 2; k < ncol; k++) Beta[(size_t)Kc * k + c] = 0;
        column(&s, c, col);
        double z = 0;
        for (int h = 0; h < N; h++) z += col[h] * col[h];
        if (z == 0) z = 1;
        s.scale[c] = sqrt(z);
    }
    free(col);
    int cap = s.cap;
    s.used = calloc(cap, sizeof(int)); s.unused = calloc(Kc, sizeof(int));
    s.alpha = calloc(cap, sizeof(double)); s.mu = calloc(cap, sizeof(double)); s.gamma = calloc(cap, sizeof(double));
    s.sigma = calloc((size_t)cap * cap, sizeof(double)); s.sigma_new = calloc((size_t)cap * cap, sizeof(double));
    s.H = calloc((size_t)cap * cap, sizeof(double));
    s.phi = calloc((size_t)N * cap, sizeof(double));
    s.G = calloc(cap, sizeof(double *));
    s.xt = calloc(Kc, sizeof(double));
    s.S_in = calloc(Kc, sizeof(double)); s.Q_in = calloc(Kc, sizeof(double));
    s.S_out = calloc(Kc, sizeof(double)); s.Q_out = calloc(Kc, sizeof(double));
    s.dml = calloc(Kc, sizeof(double)); s.aroot = calloc(Kc, sizeof(double));
    s.action = calloc(Kc, sizeof(int)); s.block = calloc(Kc, sizeof(int));
    s.qcorr = g_fresh_statistics ? calloc(Kc, sizeof(double)) : NULL;
    s.M = 1;
    double *t = malloc(sizeof(double) * N);
    double b = 0;
    for (int i = 0; i < N; i++) b += y[i];
    b /= N;
    double var_y = var_targets(y, N), eps = var_y * 0.01, residvar = 1e10;
    double vk = 1e-30, vk0, err = 1000;
    int iter = 0;
    double *w = malloc(sizeof(double) * cap), *one_phi = malloc(sizeof(double) * cap), *y_phi = malloc(sizeof(double) * cap);
    while (iter < 100 && err > 1e-8 && residvar >= eps) {      /* MainEff.c:155-197 */
        iter++;
        vk0 = vk;
        for (int i = 0; i < N; i++) t[i] = y[i] - b;
        inner_solver(&s, v, t, lambda, alpha_en, iter, residvar, var_y);
        /* b = 1'C^-1 y / (1'C^-1 1 [+1e-10]),  C^-1 = beta I - beta^2 PHI SIGMA PHI'  (:742-781, 172-188) */
        int M = s.M;
        double sy = 0;
        for (int i = 0; i < N; i++) sy += y[i];
        for (int j = 0; j < M; j++) {
            double a1 = 0, ay = 0;
            for (int h = 0; h < N; h++) { a1 += s.phi[(size_t)j * N + h]; ay += s.phi[(size_t)j * N + h] * y[h]; }
            one_phi[j] = a1; y_phi[j] = ay;
        }
        double q11 = 0, q1y = 0;
        for (int j = 0; j < M; j++) {
            double z = 0;
            for (int k = 0; k < M; k++) z += s.sigma[j * M + k] * one_phi[k];
            q11 += z * one_phi[j]; q1y += z * y_phi[j];
        }
        double cinv = s.beta * N - s.beta * s.beta * q11;
        double cinvy = s.beta * sy - s.beta * s.beta * q1y;
        b = v->epis ? cinvy / cinv : cinvy / (cinv + 1e-10);   /* MainEff.c:188 vs NeFull2.c:202 */
        vk = 0;
        for (int i = 0; i < M; i++) vk += s.alpha[i];
        err = fabs(vk - vk0) / M;
        residvar = 1 / (s.beta + 1e-10);
        for (int i = 0; i < cap; i++) if (s.G[i]) { free(s.G[i]); s.G[i] = NULL; }   /* cache is rebuilt per outer iteration (:786-790) */
    }
    int M = s.M;
    /* Wald score mu' H mu with the H left by the last FinalUpdate, read with the FINAL M as
     * leading dimension exactly as the reference does (:206-215). */
    double wd = 0;
    for (int i = 0; i < M; i++) {
        double z = 0;
        for (int j = 0; j < M; j++) z += s.mu[j] * s.H[i * M + j];
        w[i] = z;
    }
    for (int i = 0; i < M; i++) wd += w[i] * s.mu[i];
    wald[0] = wd;
    for (int i = 0; i < M; i++) {
        int c = s.used[i] - 1;
        Beta[(size_t)Kc * 2 + c] = s.mu[i] / s.scale[c];
        Beta[(size_t)Kc * 3 + c] = s.sigma[i * M + i] / (s.scale[c] * s.scale[c]);
        if (v->epis) Beta[(size_t)Kc * 4 + c] = s.used[i];
    }
    intercept[0] = b;
    residual[0] = 1 / (s.beta + 1e-10);
    free(w); free(one_phi); free(y_phi); free(t);
    free(loc1); free(loc2); free(s.scale); free(s.used); free(s.unused); free(s.alpha); free(s.mu); free(s.gamma);
    free(s.sigma); free(s.sigma_new); free(s.H); free(s.phi); free(s.G); free(s.xt);
    free(s.S_in); free(s.Q_in); free(s.S_out); free(s.Q_out); free(s.dml); free(s.aroot); free(s.action); free(s.block);
    free(s.qcorr);
}

static const Variant G_MAIN = { 0, 0.9, 1e-3, 1e-3, 1e2 };
static const Variant G_EPIS = { 1, 0.99, 1e-2, 0.1, 1e3 };

/* Same signatures as the reference `.C` entry points (MainEff.c:55-57, NeFull2.c:57-58). */
void oracle_elasticNetLinearNeMainEff(double *BASIS, double *y, double *a_lambda, double *b_Alpha, double *Beta,
                                      double *wald, double *intercept, int *n, int *kdim, int *verb, double *residual)
{
    (void)verb;
    fit(&G_MAIN, BASIS, y, *a_lambda, *b_Alpha, Beta, wald, intercept, *n, *kdim, residual);
}

void oracle_elasticNetLinearNeEpisEff(double *BASIS, double *y, double *a_lambda, double *b_Alpha, double *Beta,
                                      double *wald, double *intercept, int *n, int *kdim, int *verb, double *residual)
{
    (void)verb;
    fit(&G_EPIS, BASIS, y, *a_lambda, *b_Alpha, Beta, wald, intercept, *n, *kdim, residual);
}
