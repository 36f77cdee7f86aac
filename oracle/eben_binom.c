/* TEST INFRASTRUCTURE -- CPU restatement (oracle "port") of EBEN's Binomial (logistic)
 * empirical-Bayes elastic-net fit, main-effect and Epis variants.  Plain C, no BLAS.
 *
 * Restates (file:line under /root/reference/EBEN_orig/src/):
 *   ElasticNetBinaryNEmainEff.c  entry 236-389, inner solver 397-827, ActionAdd 830-1003,
 *                                ActionDel 1010-1121, ActionRes 1127-1203, init 1215-1406,
 *                                least squares 1588-1618, FullStat 1633-1803, PostMode (IRLS)
 *                                1808-2010, data error / sigmoid 2013-2032, DeltaML 2063-2238
 *   ElasticNetBinaryNeFull.c     same skeleton over K(K+1)/2 candidates: entry 52-226 (compact
 *                                output with locus decode 168-208), inner 234-668, init 674-804,
 *                                FullStat 848-993, PostMode 998-1152, actions 1313-1770,
 *                                DeltaML 1775-1950; constants per SURVEY.md Appendix A.
 * The 2-column least-squares start (dgelsy, rcond 1e-5) is restated as a pivoted QR with the
 * same rank decision; the SPD inverse (dpotrf/dpotri) as a plain Cholesky inverse.
 * Parity is pinned by tests/test_oracle.py against oracle/_ref (reference C compiled
 * unmodified) and tests/golden/ vectors generated from it.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may link this file.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { ACT_REEST = 0, ACT_ADD = 1, ACT_DEL = -1, ACT_TERM = 10, ACT_NONE = -10 };

typedef struct {
    int N, K, Kc, cap, epis;
    const double *X, *y;
    const int *loc1, *loc2;
    double *scale;
    int M;                 /* intercept + active effects */
    int *used, *unused; int n_unused;
    double *alpha;         /* cap entries; entries past N_used keep stale values (dasum quirk, :340) */
    double *mu, *sigma, *H; /* M, M*M (ld M) */
    double *phi;           /* N x (cap+1); column 0 = ones */
    double *w;             /* IRLS weights "beta" */
    double *S_in, *Q_in, *S_out, *Q_out, *dml, *aroot;
    int *action, *block;
    int overflow, not_pd;
} BState;

static void bcolumn(const BState *s, int c, double *out)
{
    const double *a = s->X + (size_t)s->loc1[c] * s->N;
    if (s->loc1[c] == s->loc2[c]) { memcpy(out, a, sizeof(double) * s->N); return; }
    const double *b = s->X + (size_t)s->loc2[c] * s->N;
    for (int h = 0; h < s->N; h++) out[h] = a[h] * b[h];
}

static void norm_column(const BState *s, int c, double *out)
{   /* main: dcopy + dscal by 1/scale (NEmainEff.c:643-646); Epis: element / scale (NeFull.c:253-262) */
    bcolumn(s, c, out);
    if (!s->epis) { double r = 1.0 / s->scale[c]; for (int h = 0; h < s->N; h++) out[h] *= r; }
    else for (int h = 0; h < s->N; h++) out[h] = out[h] / s->scale[c];
}

static int bspd_inverse(double *a, int n)
{   /* role of dpotrf + dpotri + mirror (NEmainEff.c:2036-2060) */
    for (int j = 0; j < n; j++) {
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
    for (int j = 0; j < n; j++) {
        a[j * n + j] = 1.0 / a[j * n + j];
        for (int i = 0; i < j; i++) {
            double v = 0;
            for (int k = i; k < j; k++) v += a[k * n + i] * a[j * n + k];
            a[j * n + i] = -v * a[j * n + j];
        }
    }
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

static void sigmoid_vec(double *y, const double *eta, int n) { for (int i = 0; i < n; i++) y[i] = 1 / (1 + exp(-eta[i])); }

/* Test switch (tests/test_oracle.py::test_binomial_irls_accept_reject_depends_on_summation_order): sum the data error
 * over the rows in DESCENDING order.  Mathematically the same number; it differs from the reference's ascending sum by
 * rounding (~1e-14), which is enough to flip the IRLS accept test `newTotalError >= errorLog` (NEmainEff.c:1991) on steps
 * whose true decrease is below that -- the fit then keeps an iterate with |g| just above 1e-6 instead of just below, and
 * its weights move by ~1e-8.  Any implementation that does not add these N terms in exactly the reference's order (a
 * parallel reduction cannot) inherits that sensitivity. */
static int g_reverse_error_sum = 0;
void oracle_set_reverse_error_sum(int on) { g_reverse_error_sum = on; }

static double data_error(double *y, const double *eta, const double *t, int n)
{   /* NEmainEff.c:2013-2025 */
    double e = 0;
    sigmoid_vec(y, eta, n);
    for (int k = 0; k < n; k++) {
        const int i = g_reverse_error_sum ? n - 1 - k : k;
        if (y[i] != 0) e = e - t[i] * log(y[i]);
        if (y[i] != 1) e = e - (1 - t[i]) * log(1 - y[i]);
    }
    return e;
}

static void linear_predictor(const BState *s, const double *mu, double *eta)
{
    int N = s->N, M = s->M;
    for (int i = 0; i < N; i++) {
        double z = 0;
        for (int j = 0; j < M; j++) z += s->phi[(size_t)j * N + i] * mu[j];
        eta[i] = z;
    }
}

static void post_mode(BState *s)
{   /* fEBCatPostMode* : Newton/IRLS to the posterior mode (NEmainEff.c:1808-2010, NeFull.c:998-1152) */
    int N = s->N, M = s->M;
    const double *t = s->y;
    double *eta = malloc(sizeof(double) * N), *y = malloc(sizeof(double) * N), *e = malloc(sizeof(double) * N);
    double *g = malloc(sizeof(double) * M), *dmu = malloc(sizeof(double) * M), *mun = malloc(sizeof(double) * M);
    double err_log[25];
    linear_predictor(s, s->mu, eta);
    double derr = data_error(y, eta, t, N), reg = 0;
    for (int i = 1; i < M; i++) reg = reg + s->alpha[i - 1] * s->mu[i] * s->mu[i] / 2;
    double total = reg + derr;
    for (int it = 0; it < 25; it++) {
        err_log[it] = total;
        double g0 = 0, h0 = 0;
        for (int j = 0; j < N; j++) {
            if (s->epis) { if (y[j] < 1e-5) y[j] = 1e-5; if (y[j] > (1 - 1e-5)) y[j] = 1 - 1e-5; }   /* NeFull.c:1049-1050 */
            e[j] = t[j] - y[j];
            g0 += e[j];
            double b = y[j] * (1 - y[j]);
            if (!s->epis) { if (b < 1e-10) b = 1e-5; if (b > 1e10) b = 1e5; }     /* NEmainEff.c:1880-1882 */
            else { if (b < 1e-5) b = 1e-3; if (b > 1e5) b = 1e3; }               /* NeFull.c:1054-1055 */
            s->w[j] = b;
            h0 += b;
        }
        g[0] = g0; s->H[0] = h0;
        for (int j = 1; j < M; j++) {
            const double *p = s->phi + (size_t)j * N;
            double gj = 0, hj = 0;
            for (int k = 0; k < N; k++) { gj += p[k] * e[k]; hj += s->w[k] * p[k]; }
            g[j] = gj - s->alpha[j - 1] * s->mu[j];
            s->H[j] = hj; s->H[j * M] = hj;
        }
        for (int j = 1; j < M; j++)
            for (int k = 1; k < M; k++) {
                const double *a = s->phi + (size_t)j * N, *b = s->phi + (size_t)k * N;
                double z = 0;
                for (int L = 0; L < N; L++) z += a[L] * s->w[L] * b[L];
                if (j == k) z += s->alpha[k - 1];
                s->H[k * M + j] = z;
            }
        memcpy(s->sigma, s->H, sizeof(double) * M * M);
        if (bspd_inverse(s->sigma, M)) s->not_pd = 1;
        int cnt = 0, need = s->epis ? M : M - 1;
        for (int j = s->epis ? 0 : 1; j < M; j++) if (fabs(g[j]) < 1e-6) cnt++;
        for (int k = 0; k < M; k++) {
            double z = 0;
            for (int L = 0; L < M; L++) z += s->sigma[L * M + k] * g[L];
            dmu[k] = z;
        }
        if (cnt == need) break;
        double step = 1;
        while (step > 1.0 / 256) {
            for (int j = 0; j < M; j++) mun[j] = s->mu[j] + step * dmu[j];
            linear_predictor(s, mun, eta);
            derr = data_error(y, eta, t, N);
            reg = 0;
            for (int j = 1; j < M; j++) reg = reg + s->alpha[j - 1] * mun[j] * mun[j] / 2;
            total = derr + reg;
            if (total >= err_log[it]) step = step / 2;
            else { memcpy(s->mu, mun, sizeof(double) * M); step = 0; }
        }
    }
    free(eta); free(y); free(e); free(g); free(dmu); free(mun);
}

static void bp_vector(const BState *s, const double *col, double sc, double *bp)
{   /* BASIS' B PHI for one candidate: bp[p] = sum_j x[j] phi_p[j] w[j] / scale (NEmainEff.c:1693-1702) */
    int N = s->N, M = s->M;
    for (int p = 0; p < M; p++) {
        const double *ph = s->phi + (size_t)p * N;
        double z = 0;
        for (int j = 0; j < N; j++) z += col[j] * ph[j] * s->w[j];
        bp[p] = z / sc;
    }
}

static void full_stat(BState *s)
{   /* fEBCatFullStat* (NEmainEff.c:1633-1803, NeFull.c:848-993) */
    int N = s->N, M = s->M;
    post_mode(s);
    double *eta = malloc(sizeof(double) * N), *y = malloc(sizeof(double) * N), *e = malloc(sizeof(double) * N);
    double *bp = malloc(sizeof(double) * M), *col = malloc(sizeof(double) * N);
    linear_predictor(s, s->mu, eta);
    sigmoid_vec(y, eta, N);
    for (int i = 0; i < N; i++) e[i] = s->y[i] - y[i];
    for (int c = 0; c < s->Kc; c++) {
        bcolumn(s, c, col);
        double sc = s->scale[c];
        bp_vector(s, col, sc, bp);
        double quad = 0;
        for (int p = 0; p < M; p++) {
            double z = 0;
            for (int j = 0; j < M; j++) z += bp[j] * s->sigma[p * M + j];
            quad += z * bp[p];
        }
        double bb = 0, ze = 0;
        for (int p = 0; p < N; p++) { bb += s->w[p] * col[p] * col[p]; ze += col[p] * e[p]; }
        s->S_in[c] = bb / (sc * sc) - quad;
        s->Q_in[c] = ze / sc;
        s->S_out[c] = s->S_in[c];
        s->Q_out[c] = s->Q_in[c];
    }
    for (int i = 0; i < M - 1; i++) {
        int c = s->used[i] - 1;
        s->S_out[c] = s->alpha[i] * s->S_in[c] / (s->alpha[i] - s->S_in[c]);
        s->Q_out[c] = s->alpha[i] * s->Q_in[c] / (s->alpha[i] - s->S_in[c]);
    }
    free(eta); free(y); free(e); free(bp); free(col);
}

typedef struct { double max; int nu; int any_delete; } BDecision;

static BDecision delta_ml(BState *s, double lambda, double alpha_en)
{   /* fEBDeltaML* (NEmainEff.c:2063-2238; NeFull.c:1775-1950) */
    const double l1 = lambda * alpha_en, l2 = lambda * (1 - alpha_en);
    int nu_used = s->M - 1, Kc = s->Kc, any_add = 0, any_del = 0, prio_add = 0, prio_del = 0;
    if (nu_used < 10) { prio_add = 1; prio_del = 0; }
    if (nu_used > 100 || (!s->epis && nu_used >= s->N)) { prio_add = 0; prio_del = 1; }
    for (int c = 0; c < Kc; c++) s->action[c] = ACT_NONE;
    double best = 0; int arg = 0;
    for (int pass = 0; pass < 2; pass++) {
        int n = pass == 0 ? nu_used : s->n_unused;
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
                    } else { s->action[c] = ACT_ADD; s->dml[c] = L; }     /* anyToAdd is never set here */
                }
            } else if (pass == 0 && nu_used > 1) {
                any_del = 1;
                s->action[c] = ACT_DEL;
                double o = s->alpha[i] - l2;
                double L = (log(o / (o + so + l2)) + pow(qo, 2) / (o + so + l2)) * 0.5 - l1 / o;
                s->dml[c] = -L;
            }
            if (s->dml[c] > best) { best = s->dml[c]; arg = c; }
        }
    }
    if ((any_add && prio_add) || (any_del && prio_del)) {
        for (int c = 0; c < Kc; c++) {
            if (s->action[c] == ACT_REEST) s->dml[c] = 0;
            else if (s->action[c] == ACT_DEL) { if (any_add && prio_add && !prio_del) s->dml[c] = 0; }
            else if (s->action[c] == ACT_ADD) { if (any_del && prio_del && !prio_add) s->dml[c] = 0; }
        }
        best = 0; arg = 0;
        for (int c = 0; c < Kc; c++) if (s->dml[c] > best) { best = s->dml[c]; arg = c; }
    }
    BDecision out = { best, arg, any_del };
    return out;
}

/* S_in/Q_in correction shared by the three actions: t_c = (BASIS'B PHI)_c . v, recomputed from
 * scratch per candidate exactly as the reference does (NEmainEff.c:967-987, 1040-1058, 1171-1194). */
static void action_add(BState *s, int nu, double new_alpha, const double *phi_new)
{   /* ActionAdd* (NEmainEff.c:830-1003) */
    int N = s->N, M = s->M, M1 = M + 1, n_used = M - 1;
    double *bphi = malloc(sizeof(double) * N), *xbphi = malloc(sizeof(double) * s->Kc);
    double *col = malloc(sizeof(double) * N), *bp = malloc(sizeof(double) * M);
    double *tmp = malloc(sizeof(double) * M), *u = malloc(sizeof(double) * M);
    double *snew = malloc(sizeof(double) * M1 * M1);
    for (int j = 0; j < N; j++) bphi[j] = s->w[j] * phi_new[j];
    for (int c = 0; c < s->Kc; c++) {
        bcolumn(s, c, col);
        double z = 0;
        for (int h = 0; h < N; h++) z += col[h] * bphi[h];
        xbphi[c] = z / s->scale[c];
    }
    for (int i = 0; i < M; i++) {
        double z = 0;
        for (int h = 0; h < N; h++) z += s->phi[(size_t)i * N + h] * bphi[h];
        tmp[i] = z;
    }
    for (int i = 0; i < M; i++) {
        double z = 0;
        for (int j = 0; j < M; j++) z += s->sigma[i * M + j] * tmp[j];
        u[i] = z;
    }
    s->alpha[n_used] = new_alpha;
    memcpy(s->phi + (size_t)M * N, phi_new, sizeof(double) * N);
    double s_ii = 1.0 / (new_alpha + s->S_in[nu]);
    double mu_i = s_ii * s->Q_in[nu];
    for (int i = 0; i < M; i++) s->mu[i] += -mu_i * u[i];
    s->mu[M] = mu_i;
    for (int i = 0; i < M; i++)
        for (int j = 0; j < M; j++) snew[j * M1 + i] = s->sigma[j * M + i] + (s_ii * u[i]) * u[j];
    for (int i = 0; i < M; i++) { snew[M * M1 + i] = -s_ii * u[i]; snew[i * M1 + M] = -s_ii * u[i]; }
    snew[M * M1 + M] = s_ii;
    /* S/Q with PHI still at M columns (the new column is at index M but the loop runs j < M) */
    for (int c = 0; c < s->Kc; c++) {
        bcolumn(s, c, col);
        bp_vector(s, col, s->scale[c], bp);
        double z = 0;
        for (int j = 0; j < M; j++) z += bp[j] * u[j];
        double mci = xbphi[c] - z;
        s->S_in[c] -= mci * mci * s_ii;
        s->Q_in[c] -= mu_i * mci;
    }
    memcpy(s->sigma, snew, sizeof(double) * M1 * M1);
    free(bphi); free(xbphi); free(col); free(bp); free(tmp); free(u); free(snew);
}

static void action_delete(BState *s, int jj)
{   /* ActionDel* (NEmainEff.c:1010-1121; NeFull.c:1523-1656) */
    int N = s->N, M = s->M, last = M - 1, j1 = jj + 1;
    double *tmp = malloc(sizeof(double) * M * M), *snew = malloc(sizeof(double) * (last * last + 1));
    double *col = malloc(sizeof(double) * N), *bp = malloc(sizeof(double) * M);
    const double *sj = s->sigma + j1 * M;
    double sjj = sj[j1];
    double mujj = s->mu[j1];
    for (int i = 0; i < M; i++) s->mu[i] = s->mu[i] - mujj * sj[i] / sjj;
    for (int c = 0; c < s->Kc; c++) {
        bcolumn(s, c, col);
        bp_vector(s, col, s->scale[c], bp);
        double z = 0;
        for (int j = 0; j < M; j++) z += bp[j] * sj[j];
        s->S_in[c] += pow(z, 2) / sjj;
        s->Q_in[c] += z * mujj / sjj;
    }
    for (int i = 0; i < M; i++)
        for (int j = 0; j < M; j++)
            tmp[j * M + i] = s->epis ? s->sigma[j * M + i] - sj[i] / sjj * sj[j]      /* NeFull.c:1612 */
                                     : s->sigma[j * M + i] - sj[i] * sj[j] / sjj;     /* NEmainEff.c:1069 */
    for (int i = 0; i < last; i++)
        for (int j = 0; j < last; j++) snew[j * last + i] = tmp[j * M + i];
    if (j1 != last) {
        s->alpha[jj] = s->alpha[last - 1];
        s->mu[j1] = s->mu[last];
        memmove(s->phi + (size_t)j1 * N, s->phi + (size_t)last * N, sizeof(double) * N);
        for (int i = 0; i < last; i++) snew[j1 * last + i] = tmp[last * M + i];
        tmp[j1 * M + M - 1] = tmp[M * M - 1];
        for (int i = 0; i < last; i++) snew[i * last + j1] = tmp[i * M + M - 1];
    }
    memcpy(s->sigma, snew, sizeof(double) * last * last);
    free(tmp); free(snew); free(col); free(bp);
}

static void action_reestimate(BState *s, int jj, double new_alpha)
{   /* ActionRes* (NEmainEff.c:1127-1203) */
    int N = s->N, M = s->M, j1 = jj + 1;
    double *snew = malloc(sizeof(double) * M * M), *col = malloc(sizeof(double) * N), *bp = malloc(sizeof(double) * M);
    double old = s->alpha[jj];
    s->alpha[jj] = new_alpha;
    double kappa = 1.0 / (s->sigma[j1 * M + j1] + 1.0 / (new_alpha - old));
    double mujj = s->mu[j1];
    double *sj = malloc(sizeof(double) * M);
    memcpy(sj, s->sigma + j1 * M, sizeof(double) * M);
    for (int i = 0; i < M; i++) s->mu[i] += -mujj * kappa * sj[i];
    for (int i = 0; i < M; i++)
        for (int j = 0; j < M; j++) snew[j * M + i] = s->sigma[j * M + i] - kappa * sj[i] * sj[j];
    memcpy(s->sigma, snew, sizeof(double) * M * M);
    /* the reference reads SIGMA[j1*M + j] AFTER the copy, i.e. the updated row (:1166, 1191) */
    const double *sj_new = s->sigma + j1 * M;
    for (int c = 0; c < s->Kc; c++) {
        bcolumn(s, c, col);
        bp_vector(s, col, s->scale[c], bp);
        double z = 0;
        for (int j = 0; j < M; j++) z += bp[j] * sj_new[j];
        s->S_in[c] += pow(z, 2) * kappa;
        s->Q_in[c] += mujj * kappa * z;
    }
    free(snew); free(col); free(bp); free(sj);
}

static void initialise(BState *s)
{   /* fEBInitialization* (NEmainEff.c:1239-1385).  The first basis is always candidate 1. */
    int N = s->N;
    s->M = 2;
    s->used[0] = 1;
    for (int i = 0; i < N; i++) s->phi[i] = 1;
    norm_column(s, 0, s->phi + N);
    /* least squares of [1 phi] on pseudo-logits (dgelsy, rcond = 1e-5, :1366-1370) */
    double *z = malloc(sizeof(double) * N);
    for (int i = 0; i < N; i++) { double tp = 2 * s->y[i] - 1; double pp = (tp * 0.9 + 1) / 2; z[i] = log(pp / (1 - pp)); }
    const double *ph = s->phi + N;
    double n0 = sqrt((double)N), n1 = 0;
    for (int i = 0; i < N; i++) n1 += ph[i] * ph[i];
    n1 = sqrt(n1);
    double w0, w1;
    {
        /* pivoted QR: the larger-norm column leads */
        int lead_ones = n0 >= n1;
        double r11 = lead_ones ? n0 : n1, r12 = 0, c1 = 0;
        double *q1 = malloc(sizeof(double) * N), *v = malloc(sizeof(double) * N);
        for (int i = 0; i < N; i++) q1[i] = (lead_ones ? 1.0 : ph[i]) / r11;
        for (int i = 0; i < N; i++) { double other = lead_ones ? ph[i] : 1.0; r12 += q1[i] * other; c1 += q1[i] * z[i]; }
        double r22 = 0, c2 = 0;
        for (int i = 0; i < N; i++) { double other = lead_ones ? ph[i] : 1.0; v[i] = other - r12 * q1[i]; r22 += v[i] * v[i]; }
        r22 = sqrt(r22);
        /* singular values of [[r11 r12],[0 r22]] for the rank test smax*rcond <= smin */
        double f = r11 * r11 + r12 * r12 + r22 * r22, dd = r11 * r22;
        double disc = sqrt(fmax(f * f - 4 * dd * dd, 0.0));
        double smax = sqrt((f + disc) / 2), smin = smax > 0 ? fabs(dd) / smax : 0;
        double xa, xb;      /* coefficients of (leading, trailing) column */
        if (r22 > 0 && smax * 1e-5 <= smin) {
            for (int i = 0; i < N; i++) c2 += v[i] * z[i];
            xb = c2 / (r22 * r22);
            xa = (c1 - r12 * xb) / r11;
        } else {            /* rank 1: minimum-norm solution of [r11 r12] x = c1 */
            double den = r11 * r11 + r12 * r12;
            xa = r11 * c1 / den; xb = r12 * c1 / den;
        }
        if (lead_ones) { w0 = xa; w1 = xb; } else { w0 = xb; w1 = xa; }
        free(q1); free(v);
    }
    free(z);
    s->mu[0] = w0; s->mu[1] = w1;
    double a = w1 == 0 ? 1 : 1 / (w1 * w1);
    if (a < 1e-3) a = 1e-3;
    if (a > 1e3) a = 1e3;
    s->alpha[0] = a;
}

static void rebuild_unused(BState *s)
{
    char *isused = calloc(s->Kc + 1, 1);
    for (int j = 0; j < s->M - 1; j++) isused[s->used[j]] = 1;
    int kk = 0;
    for (int c = 0; c < s->Kc; c++) if (!isused[c + 1]) s->unused[kk++] = c + 1;
    s->n_unused = kk;
    free(isused);
}

static double inner_solver(BState *s, double lambda, double alpha_en, int iter, double n_add)
{   /* fEBBinaryMex* (NEmainEff.c:397-827) */
    int N = s->N, Kc = s->Kc;
    int ini_removed = 1;
    if (iter <= 1) { initialise(s); ini_removed = 0; }
    rebuild_unused(s);
    int initial = s->used[0];
    full_stat(s);
    int selected = ACT_NONE, last = 0, n_update = 0, jj = -1, i_iter = 0;
    int it_max = iter == 1 ? 10 : 100;
    double loglik = 1e-30;
    double *phi_new = malloc(sizeof(double) * N), *eta = malloc(sizeof(double) * N);
    while (!last) {
        i_iter++;
        double logl0 = loglik;
        BDecision d = delta_ml(s, lambda, alpha_en);
        int nu = d.nu, worthwhile;
        if (selected == ACT_TERM && !ini_removed && s->M > 2) nu = -1;
        if (nu == -1 && ini_removed) { worthwhile = 0; selected = ACT_TERM; }
        else if (nu == -1 && !ini_removed && s->M > 2) {
            worthwhile = 1; nu = initial - 1;
            s->action[nu] = ACT_DEL; n_update = 1; s->block[0] = nu; ini_removed = 1; selected = ACT_DEL;
        } else {
            worthwhile = 1;
            double cutoff = d.max * (s->action[nu] == ACT_ADD ? n_add : 1.0);
            if (cutoff < 1e-3) cutoff = 1e-3;
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
                    for (int i = 0; i < s->M - 1; i++) if (s->used[i] == nu + 1) jj = i;
                norm_column(s, nu, phi_new);
                if (selected == ACT_REEST && fabs(log(new_alpha) - log(s->alpha[jj])) <= 1e-3 && !d.any_delete)
                    selected = ACT_TERM;
                if (selected == ACT_REEST) action_reestimate(s, jj, new_alpha);
                else if (selected == ACT_ADD) {
                    int n_used = s->M - 1;
                    if (n_used + 1 > s->cap) { s->overflow = 1; selected = ACT_TERM; }  /* reference: printf + return / overrun (:692-695) */
                    else {
                        action_add(s, nu, new_alpha, phi_new);
                        s->used[n_used] = nu + 1;
                        s->n_unused--;
                        for (int i = 0; i < s->n_unused; i++) if (s->unused[i] == nu + 1) s->unused[i] = s->unused[s->n_unused];
                        s->M++;
                    }
                } else if (selected == ACT_DEL) {
                    action_delete(s, jj);
                    int idx = s->M - 2;
                    s->used[jj] = s->used[idx];
                    s->unused[s->n_unused++] = nu + 1;
                    s->M--;
                    if (nu + 1 == initial) ini_removed = 1;                  /* :744 */
                }
                if (iu == n_update - 1) full_stat(s);                        /* :749-762 */
            }
        }
        if (selected == ACT_TERM && ini_removed) last = 1;
        if ((i_iter == it_max && s->M == 2) || i_iter > it_max) last = 1;
        if (i_iter == it_max) selected = ACT_TERM;
        linear_predictor(s, s->mu, eta);
        loglik = 0;
        for (int i = 0; i < N; i++)
            loglik = loglik + s->y[i] * log(exp(eta[i]) / (1 + exp(eta[i]))) + (1 - s->y[i]) * log(1 / (1 + exp(eta[i])));
        double dL = fabs((loglik - logl0) / logl0);
        if (dL < 1e-3) selected = ACT_TERM;
    }
    free(phi_new); free(eta);
    return loglik;
}

static void fit(int epis, const double *X, const double *y, double lambda, double alpha_en, double *logl_out,
                double *Beta, double *wald, double *intercept, int N, int K, int bmax)
{
    BState s; memset(&s, 0, sizeof s);
    int Kc = epis ? (K + 1) * K / 2 : K;
    s.N = N; s.K = K; s.Kc = Kc; s.X = X; s.y = y; s.epis = epis; s.cap = bmax;
    int *loc1 = malloc(sizeof(int) * Kc), *loc2 = malloc(sizeof(int) * Kc);
    for (int i = 0; i < K; i++) loc1[i] = loc2[i] = i;
    if (epis) { int kk = K; for (int i = 0; i < K - 1; i++) for (int j = i + 1; j < K; j++) { loc1[kk] = i; loc2[kk] = j; kk++; } }
    s.loc1 = loc1; s.loc2 = loc2;
    s.scale = malloc(sizeof(double) * Kc);
    double *col = malloc(sizeof(double) * N);
    for (int c = 0; c < Kc; c++) {
        bcolumn(&s, c, col);
        double z = 0;
        for (int h = 0; h < N; h++) z += col[h] * col[h];
        if (z == 0) z = 1;
        s.scale[c] = sqrt(z);
    }
    free(col);
    if (!epis) {   /* NEmainEff.c:270-291 */
        for (int i = 0; i < K; i++) { Beta[i] = i + 1; Beta[K + i] = i + 1; }
        for (int i = 0; i < bmax; i++) { Beta[bmax * 2 + i] = 0; Beta[bmax * 3 + i] = 0; }
    } else for (int i = 0; i < bmax; i++) Beta[bmax * 2 + i] = 0;      /* NeFull.c:67 */
    int cap1 = bmax + 2;
    s.used = calloc(cap1, sizeof(int)); s.unused = calloc(Kc, sizeof(int));
    s.alpha = calloc(cap1, sizeof(double)); s.mu = calloc(cap1, sizeof(double));
    s.sigma = calloc((size_t)cap1 * cap1, sizeof(double)); s.H = calloc((size_t)cap1 * cap1, sizeof(double));
    s.phi = calloc((size_t)N * cap1, sizeof(double)); s.w = calloc(N, sizeof(double));
    s.S_in = calloc(Kc, sizeof(double)); s.Q_in = calloc(Kc, sizeof(double));
    s.S_out = calloc(Kc, sizeof(double)); s.Q_out = calloc(Kc, sizeof(double));
    s.dml = calloc(Kc, sizeof(double)); s.aroot = calloc(Kc, sizeof(double));
    s.action = calloc(Kc, sizeof(int)); s.block = calloc(Kc, sizeof(int));
    s.M = 2;
    double vk = 1e-30, vk0, err = 1000, loglik = 0;
    int iter = 0;
    while (iter < 100 && err > 1e-8) {          /* NEmainEff.c:329-344 */
        iter++;
        vk0 = vk;
        loglik = inner_solver(&s, lambda, alpha_en, iter, epis ? 0.99 : 0.90);
        vk = 0;
        int upto = epis ? s.M - 1 : s.M;          /* main: dasum over m entries (one stale slot, :340); Epis: m-1 (:139) */
        for (int i = 0; i < upto; i++) vk += epis ? s.alpha[i] : fabs(s.alpha[i]);
        err = fabs(vk - vk0) / s.M;
    }
    int M = s.M;
    double wd = 0;
    for (int i = 0; i < M; i++) {
        double z = 0;
        for (int j = 0; j < M; j++) z += s.mu[j] * s.H[i * M + j];
        wd += z * s.mu[i];
    }
    wald[0] = wd;
    if (!epis) {
        for (int i = 1; i < M; i++) {
            int c = s.used[i - 1] - 1;
            Beta[bmax * 2 + c] = s.mu[i] / s.scale[c];
            Beta[bmax * 3 + c] = s.sigma[i * M + i] / (s.scale[c] * s.scale[c]);
        }
    } else {
        int meff = M - 1;
        if (M > bmax) meff = bmax;
        for (int i = 0; i < meff; i++) {
            int c = s.used[i] - 1;
            Beta[i] = loc1[c] + 1; Beta[bmax + i] = loc2[c] + 1;
            Beta[bmax * 2 + i] = s.mu[i + 1] / s.scale[c];
            Beta[bmax * 3 + i] = s.sigma[(i + 1) * M + i + 1] / (s.scale[c] * s.scale[c]);
        }
    }
    intercept[0] = s.mu[0]; intercept[1] = s.sigma[0];
    logl_out[0] = loglik;
    free(loc1); free(loc2); free(s.scale); free(s.used); free(s.unused); free(s.alpha); free(s.mu);
    free(s.sigma); free(s.H); free(s.phi); free(s.w); free(s.S_in); free(s.Q_in); free(s.S_out); free(s.Q_out);
    free(s.dml); free(s.aroot); free(s.action); free(s.block);
}

/* Same signatures as the reference `.C` entry points (NEmainEff.c:236-238, NeFull.c:52-55). */
void oracle_ElasticNetBinaryNEmainEff(double *BASIS, double *Targets, double *a_Lambda, double *b_Alpha,
                                      double *logLIKELIHOOD, double *Beta, double *wald, double *intercept,
                                      int *n, int *kdim, int *VB, int *bMax)
{
    (void)VB;
    fit(0, BASIS, Targets, *a_Lambda, *b_Alpha, logLIKELIHOOD, Beta, wald, intercept, *n, *kdim, *bMax);
}

void oracle_ElasticNetBinaryNEfull(double *BASIS, double *Targets, double *a_Lambda, double *b_Alpha,
                                   double *logLIKELIHOOD, double *Beta, double *wald, double *intercept,
                                   int *n, int *kdim, int *VB, int *bMax)
{
    (void)VB;
    fit(1, BASIS, Targets, *a_Lambda, *b_Alpha, logLIKELIHOOD, Beta, wald, intercept, *n, *kdim, *bMax);
}
