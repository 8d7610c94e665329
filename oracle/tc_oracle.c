/*
 * tc_oracle.c — plain-C (FP64) CPU restatement of the reference's hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): only tests/, smoke() and
 * bench.py's cpu_baseline / --impl reference legs may link or call this; the product
 * library (libtcmcmc.so) never does.  It doubles as the timed CPU baseline ("port").
 *
 * What it follows (paths relative to /root/reference):
 *   orc_colon / orc_t_interp   MATLAB colon semantics used at
 *                              src/SumofSquaresFunction_TranscriptionCycleMCMC.m:29-30
 *   orc_elongation_sim         src/dependencies/ConstantElongationSim.m:1-69 (m x n matrix)
 *   orc_fluor                  src/GetFluorFromPolPos.m:1-71
 *   orc_model_on_grid          SumofSquares...m:49-51 / src/TranscriptionCycleMCMC.m:307-309
 *   orc_ss                     src/SumofSquaresFunction_TranscriptionCycleMCMC.m:1-65
 *   orc_dram                   mcmcstat::mcmcrun's DRAM loop as configured at
 *                              src/TranscriptionCycleMCMC.m:242-273.  mcmcstat is a third-party
 *                              package, NOT vendored / version-pinned by the reference
 *                              (README.md:5): restated from the published algorithm
 *                              (Haario, Laine, Mira, Saksman 2006) + its documented defaults.
 *                              PARITY UNPINNED beyond what the 10-step fixture pins
 *                              (SURVEY.md 4.3).
 *
 * The literal m x n algorithm is kept on purpose (no cohort shortcut): it is the oracle
 * the CUDA cohort/Toeplitz formulation is validated against, and the same O(m*n) work the
 * MATLAB reference does.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_SETS 8

typedef struct {
    int nsets;
    double L_ms2, L_pp7;                 /* base lengths; tau*v is added (GetFluorFromPolPos.m:19-20) */
    double ms2_start[ORC_MAX_SETS], ms2_end[ORC_MAX_SETS], ms2_loopn[ORC_MAX_SETS];
    double pp7_start[ORC_MAX_SETS], pp7_end[ORC_MAX_SETS], pp7_loopn[ORC_MAX_SETS];
} orc_construct;

typedef struct {
    int nsimu;          /* options.nsimu = n_steps                 (TranscriptionCycleMCMC.m:264) */
    int burnintime;     /* options.burnintime = n_burn             (:267) */
    int adaptint;       /* options.adaptint = 100                  (:268) */
    int ntry;           /* 'dram' => 2 (one delayed-rejection retry) [fixture: DR scale 5.007] */
    int updatesigma;    /* options.updatesigma = 1                 (:265) */
    int burnin_cumulative; /* 1 (default): cumulative rejection rate, mcmcstat's `rejected > 0.95*isimu` [U: SURVEY 3.2 / B.3];
                              0: rejection rate since the last adaptation */
    int qcovadj_always; /* 0 (default): chol(cov) first, chol(cov + qcovadj*I) only when that fails — mcmcstat's
                           "try to blow it" branch [U]; 1: always factor cov + qcovadj*I */
    double drscale;     /* 5   [fixture-pinned] */
    double adascale;    /* <=0 => 2.4/sqrt(npar) [mcmcstat default, unpinned] */
    double qcovadj;     /* 1e-8 [mcmcstat default, unpinned] */
    double burnin_scale;/* 10  [mcmcstat default, unpinned] */
    double N0;          /* 1   [unpinned; fixture cannot separate 0 from 1] */
    double S20;         /* = sigma2_0 = 1 */
    double sigma2_0;    /* model.sigma2 = 1                        (:212,259) */
    double Nobs;        /* model.N = length(ydata) = 2*N incl. NaNs (:260) */
} orc_dram_opts;

/* ------------------------------------------------------------------ MATLAB built-ins */

static double ml_round(double x) { return x >= 0 ? floor(x + 0.5) : -floor(-x + 0.5); }

/* a:d:b, general (non-integer) case + integer cases; returns count, fills out (cap elements) */
int orc_colon(double a, double d, double b, double *out, int cap)
{
    if (d == 0 || (d > 0 && a > b) || (d < 0 && a < b) || isnan(a) || isnan(b) || isnan(d)) return 0;
    const double eps = 2.220446049250313e-16;
    double tol = 2.0 * eps * fmax(fabs(a), fabs(b));
    double sig = d > 0 ? 1.0 : -1.0;
    long n;
    if (a == floor(a) && d == 1) {
        n = (long)(floor(b) - a);
    } else if (a == floor(a) && d == floor(d)) {
        double q = floor(a / d), r = a - q * d;
        n = (long)(floor((b - r) / d) - q);
    } else {
        n = (long)ml_round((b - a) / d);
        if (sig * (a + n * d - b) > tol) n -= 1;
    }
    double c = a + n * d;
    if (sig * (c - b) > -tol) c = b;
    if (n + 1 > cap) return -(int)(n + 1);
    for (long k = 0; k <= n / 2; ++k) {
        out[k] = a + k * d;
        out[n - k] = c - k * d;
    }
    if (n % 2 == 0) out[n / 2] = (a + c) / 2;
    return (int)(n + 1);
}

/* dt = mean(t(2:end)-t(1:end-1)); t_interp = t(1):dt:t(end)   (SumofSquares...m:29-30) */
int orc_t_interp(int N, const double *t, double *ti)
{
    double s = 0;
    for (int i = 0; i + 1 < N; ++i) s += t[i + 1] - t[i];
    double dt = s / (N - 1);
    return orc_colon(t[0], dt, t[N - 1], ti, N);
}

/* ------------------------------------------------------------------ forward model */

/* ConstantElongationSim.m: returns malloc'd m x n row-major matrix, *n_out = n */
static double *orc_elongation_sim(double v, double ton, const double *Rfull, int m, const double *t,
                                  int *n_out)
{
    double *R = (double *)malloc(sizeof(double) * (m > 1 ? m - 1 : 1));
    double *dt = (double *)malloc(sizeof(double) * (m > 1 ? m - 1 : 1));
    double tot = 0;
    for (int i = 0; i < m - 1; ++i) {
        R[i] = Rfull[i] < 0 ? 0.0 : Rfull[i];          /* :33,:36 */
        dt[i] = t[i + 1] - t[i];                        /* :43-45 */
        tot += R[i] * dt[i];
    }
    long n = (long)floor(tot);                          /* :47 */
    if (n < 0) n = 0;
    /* the running counter can exceed sum() by rounding only; MATLAB would grow the matrix */
    long ncap = n + 2;
    double *x = (double *)calloc((size_t)m * (size_t)ncap, sizeof(double)); /* :50 */
    double counter = 0;                                 /* :53 */
    for (int i = 0; i < m - 1; ++i) {                   /* :56 */
        if (t[i] < ton) continue;                       /* :57-58 */
        counter += R[i] * dt[i];                        /* :60 */
        long k = (long)floor(counter);                  /* :61 */
        if (k > ncap) k = ncap;
        const double *xi = x + (size_t)i * ncap;
        double *xn = x + (size_t)(i + 1) * ncap;
        double step = v * dt[i];
        for (long c = 0; c < k; ++c) xn[c] = xi[c] + step;   /* :64 */
        /* :65 is a no-op for v >= 0 */
    }
    free(R);
    free(dt);
    *n_out = (int)ncap;
    return x;
}

/* GetFluorFromPolPos.m:47-70 on an m x n matrix */
static void orc_fluor(const orc_construct *c, const double *x, int m, int n, double v, double tau,
                      double b_ms2, double b_pp7, double *MS2, double *PP7)
{
    double L_ms2 = c->L_ms2 + tau * v, L_pp7 = c->L_pp7 + tau * v;   /* :19-20 */
    for (int j = 0; j < m; ++j) { MS2[j] = 0; PP7[j] = 0; }           /* :29-30 */
    for (int s = 0; s < c->nsets; ++s) {                               /* :47 */
        double fv = c->ms2_loopn[s] / 24, st = c->ms2_start[s], en = c->ms2_end[s];
        for (int j = 0; j < m; ++j) {
            const double *row = x + (size_t)j * n;
            double acc = 0;
            for (int k = 0; k < n; ++k) {
                double p = row[k], val = 0;
                if (p > en && p < L_ms2) val = fv;                      /* :50 */
                if (p > st && p < en) val = (p - st) * fv / (en - st); /* :51-52 */
                acc += val;
            }
            MS2[j] += acc;                                              /* :54 */
            if (MS2[j] < b_ms2) MS2[j] = b_ms2;                         /* :57 */
        }
        fv = c->pp7_loopn[s] / 24; st = c->pp7_start[s]; en = c->pp7_end[s];
        for (int j = 0; j < m; ++j) {
            const double *row = x + (size_t)j * n;
            double acc = 0;
            for (int k = 0; k < n; ++k) {
                double p = row[k], val = 0;
                if (p > en && p < L_pp7) val = fv;                      /* :62 */
                if (p > st && p < en) val = (p - st) * fv / (en - st); /* :63-64 */
                acc += val;
            }
            PP7[j] += acc;                                              /* :66 */
            if (PP7[j] < b_pp7) PP7[j] = b_pp7;                         /* :69 */
        }
    }
}

/* theta = [v,tau,ton,MS2_basal,PP7_basal,A,R,dR_1..dR_m]; outputs A*MS2 and PP7 on tgrid */
void orc_model_on_grid(const orc_construct *c, int m, const double *tgrid, const double *theta,
                       double *ms2, double *pp7)
{
    double v = theta[0], tau = theta[1], ton = theta[2], b1 = theta[3], b2 = theta[4], A = theta[5],
           R = theta[6];
    double *Rf = (double *)malloc(sizeof(double) * m);
    for (int i = 0; i < m; ++i) Rf[i] = R + theta[7 + i];               /* SumofSquares...m:45 */
    int n;
    double *x = orc_elongation_sim(v, ton, Rf, m, tgrid, &n);
    orc_fluor(c, x, m, n, v, tau, b1, b2, ms2, pp7);
    for (int j = 0; j < m; ++j) ms2[j] = A * ms2[j];                     /* :51 */
    free(x);
    free(Rf);
}

/* interp1 linear with NaN outside the grid (SumofSquares...m:55-56) */
static double interp1_lin(int n, const double *x, const double *v, double z)
{
    if (!(z >= x[0] && z <= x[n - 1])) return NAN;
    int lo = 0, hi = n - 1;               /* find k: x[k] <= z < x[k+1] */
    while (hi - lo > 1) {
        int mid = (lo + hi) / 2;
        if (x[mid] <= z) lo = mid; else hi = mid;
    }
    double s = (z - x[lo]) / (x[lo + 1] - x[lo]);
    return v[lo] + s * (v[lo + 1] - v[lo]);
}

double orc_ss(const orc_construct *c, int N, const double *t, const double *ms2e, const double *pp7e,
              const double *theta)
{
    double *ti = (double *)malloc(sizeof(double) * 3 * N);
    double *m1 = ti + N, *m2 = ti + 2 * N;
    int cnt = orc_t_interp(N, t, ti);
    if (cnt != N) { free(ti); return NAN; }   /* MATLAB: R.*dt dimension error */
    orc_model_on_grid(c, N, ti, theta, m1, m2);
    double ss = 0;
    for (int j = 0; j < N; ++j) {              /* residuals of [MS2, PP7], nansum  :57-64 */
        double r = ms2e[j] - interp1_lin(N, ti, m1, t[j]);
        if (!isnan(r)) ss += r * r;
    }
    for (int j = 0; j < N; ++j) {
        double r = pp7e[j] - interp1_lin(N, ti, m2, t[j]);
        if (!isnan(r)) ss += r * r;
    }
    free(ti);
    return ss;
}

void orc_ss_batch(const orc_construct *c, const int *N, const long long *off, const double *t,
                  const double *ms2, const double *pp7, long long nbatch, const int *cell_id,
                  const double *theta, int ld, double *ss_out, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 16)
    for (long long b = 0; b < nbatch; ++b) {
        int cid = cell_id[b];
        ss_out[b] = orc_ss(c, N[cid], t + off[cid], ms2 + off[cid], pp7 + off[cid], theta + b * (long long)ld);
    }
}

/* ------------------------------------------------------------------ RNG for stand-alone runs */

typedef struct { uint64_t s[4]; int have; double spare; } orc_rng;
static uint64_t splitmix(uint64_t *x) { uint64_t z = (*x += 0x9e3779b97f4a7c15ULL); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; return z ^ (z >> 31); }
static void rng_seed(orc_rng *r, uint64_t seed) { for (int i = 0; i < 4; ++i) r->s[i] = splitmix(&seed); r->have = 0; }
static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t rng_next(orc_rng *r) { uint64_t *s = r->s, res = rotl(s[0] + s[3], 23) + s[0], t = s[1] << 17; s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45); return res; }
static double rng_unif(orc_rng *r) { return ((rng_next(r) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
static double rng_norm(orc_rng *r)
{
    if (r->have) { r->have = 0; return r->spare; }
    double u, v, s;
    do { u = 2 * rng_unif(r) - 1; v = 2 * rng_unif(r) - 1; s = u * u + v * v; } while (s >= 1 || s == 0);
    double f = sqrt(-2 * log(s) / s);
    r->spare = v * f; r->have = 1;
    return u * f;
}
static double rng_gamma(orc_rng *r, double a)   /* Marsaglia-Tsang, a >= 1 */
{
    double d = a - 1.0 / 3, c = 1 / sqrt(9 * d);
    for (;;) {
        double x, v;
        do { x = rng_norm(r); v = 1 + c * x; } while (v <= 0);
        v = v * v * v;
        double u = rng_unif(r);
        if (u < 1 - 0.0331 * x * x * x * x) return d * v;
        if (log(u) < 0.5 * x * x + d * (1 - v + log(v))) return d * v;
    }
}
static double rng_chi2(orc_rng *r, double nu) { return 2 * rng_gamma(r, 0.5 * nu); }

/* ------------------------------------------------------------------ DRAM */

/* prior "sum of squares": sum(((th-mu)./sig).^2); sig = Inf contributes 0 */
static double prior_ss(int npar, const double *th, const double *mu, const double *sig)
{
    double s = 0;
    for (int i = 0; i < npar; ++i) { double e = (th[i] - mu[i]) / sig[i]; s += e * e; }
    return s;
}
static int out_of_bounds(int npar, const double *th, const double *lo, const double *hi)
{
    for (int i = 0; i < npar; ++i) if (th[i] < lo[i] || th[i] > hi[i]) return 1;
    return 0;
}
/* y = x + z*R, R upper-triangular row-major npar x npar */
static void propose(int npar, const double *x, const double *z, const double *R, double scale, double *y)
{
    for (int j = 0; j < npar; ++j) {
        double s = 0;
        for (int i = 0; i <= j; ++i) s += z[i] * R[(size_t)i * npar + j];
        y[j] = x[j] + s * scale;
    }
}
/* ||d * inv(R)||^2 : solve y*R = d by forward substitution over columns */
static double norm2_dinvR(int npar, const double *d, const double *R, double *y)
{
    double nn = 0;
    for (int j = 0; j < npar; ++j) {
        double s = d[j];
        for (int i = 0; i < j; ++i) s -= y[i] * R[(size_t)i * npar + j];
        y[j] = s / R[(size_t)j * npar + j];
        nn += y[j] * y[j];
    }
    return nn;
}
/* upper Cholesky R'R = A (row-major, upper part of A used); returns 0 ok, 1 not PD */
static int chol_upper(int n, const double *A, double *R)
{
    memset(R, 0, sizeof(double) * (size_t)n * n);
    for (int j = 0; j < n; ++j) {
        for (int i = 0; i <= j; ++i) {
            double s = A[(size_t)i * n + j];
            for (int k = 0; k < i; ++k) s -= R[(size_t)k * n + i] * R[(size_t)k * n + j];
            if (i == j) {
                if (!(s > 0)) return 1;
                R[(size_t)j * n + j] = sqrt(s);
            } else {
                R[(size_t)i * n + j] = s / R[(size_t)i * n + i];
            }
        }
    }
    return 0;
}

/*
 * One chain.  Row 0 of chain/s2chain is x0 / sigma2_0; rows k = 1..nsimu-1 are MATLAB's
 * isimu = k+1.  Randomness: when z1 != NULL the five streams are consumed (row k for step k):
 * z1,z2 [nsimu x npar], u1,u2,chi2 [nsimu]; otherwise the internal generator (seed) is used and,
 * if rec_* are non-NULL, the draws that were made are recorded there (unused slots = NaN).
 * flags[k]: bit0 accepted, bit1 accepted at stage 2, bit2 stage-1 out of bounds,
 *           bit3 DR attempted, bit4 stage-2 out of bounds.
 * counters: [0] ss evaluations, [1] stage-1 accepts, [2] stage-2 accepts, [3] out-of-bound
 *           proposals, [4] adaptations done, [5] cholesky failures.
 * Returns 0, or 1 if ss(x0) is not finite.
 */
int orc_dram(const orc_construct *c, int N, const double *t, const double *ms2, const double *pp7,
             const orc_dram_opts *o, const double *theta0, const double *qcov_diag, const double *low,
             const double *upp, const double *pmu, const double *psig,
             const double *z1s, const double *u1s, const double *z2s, const double *u2s,
             const double *chi2s, uint64_t seed,
             double *rec_z1, double *rec_u1, double *rec_z2, double *rec_u2, double *rec_chi2,
             double *chain, double *s2chain, double *sschain, int *flags, long long *counters)
{
    const int npar = 7 + N, nsimu = o->nsimu;
    const double adascale = o->adascale > 0 ? o->adascale : 2.4 / sqrt((double)npar);
    const size_t np2 = (size_t)npar * npar;
    double *R = (double *)calloc(np2, sizeof(double));
    double *cov = (double *)calloc(np2, sizeof(double));
    double *tmpA = (double *)calloc(np2, sizeof(double));
    double *Rnew = (double *)calloc(np2, sizeof(double));
    double *cmean = (double *)calloc(npar, sizeof(double));
    double *old = (double *)malloc(sizeof(double) * npar * 8);
    double *y1 = old + npar, *y2 = old + 2 * npar, *zb = old + 3 * npar, *dd = old + 4 * npar,
           *wk = old + 5 * npar, *scr = old + 6 * npar, *ybuf = old + 7 * npar;
    orc_rng rng; rng_seed(&rng, seed);
    for (int i = 0; i < 6; ++i) counters[i] = 0;
    for (int i = 0; i < npar; ++i) R[(size_t)i * npar + i] = sqrt(qcov_diag[i]);  /* chol(diag(J0)) */
    memcpy(old, theta0, sizeof(double) * npar);
    double ss = orc_ss(c, N, t, ms2, pp7, old); counters[0]++;
    double pri = prior_ss(npar, old, pmu, psig);
    double sigma2 = o->sigma2_0;
    memcpy(chain, old, sizeof(double) * npar);
    s2chain[0] = sigma2;
    if (sschain) sschain[0] = ss;
    if (flags) flags[0] = 0;
    if (!isfinite(ss)) { free(R); free(cov); free(tmpA); free(Rnew); free(cmean); free(old); return 1; }
    double wsum = 0; int have_cov = 0; int lasti = 0;   /* rows [lasti, k] pending for covupd */
    long long rej = 0, reju = 0;

    for (int k = 1; k < nsimu; ++k) {
        const int isimu = k + 1;
        int accept = 0, fl = 0;
        /* ---- stage 1 */
        const double *z1;
        if (z1s) z1 = z1s + (size_t)k * npar;
        else { for (int i = 0; i < npar; ++i) zb[i] = rng_norm(&rng); z1 = zb; }
        if (rec_z1) memcpy(rec_z1 + (size_t)k * npar, z1, sizeof(double) * npar);
        propose(npar, old, z1, R, 1.0, y1);
        double ss1, pri1, a12;
        if (out_of_bounds(npar, y1, low, upp)) {
            ss1 = INFINITY; pri1 = 0; a12 = 0; fl |= 4; counters[3]++;
        } else {
            ss1 = orc_ss(c, N, t, ms2, pp7, y1); counters[0]++;
            pri1 = prior_ss(npar, y1, pmu, psig);
            a12 = exp(-0.5 * ((ss1 - ss) / sigma2 + pri1 - pri));
            if (a12 <= 0) accept = 0;
            else if (a12 >= 1) accept = 1;
            else {
                double u = u1s ? u1s[k] : rng_unif(&rng);
                if (rec_u1) rec_u1[k] = u;
                accept = a12 > u;
            }
            if (accept) counters[1]++;
        }
        const double *newp = y1; double ssn = ss1, prin = pri1;
        /* ---- delayed rejection, one retry with R/drscale */
        if (!accept && o->ntry >= 2) {
            fl |= 8;
            const double *z2;
            if (z2s) z2 = z2s + (size_t)k * npar;
            else { for (int i = 0; i < npar; ++i) wk[i] = rng_norm(&rng); z2 = wk; }
            if (rec_z2) memcpy(rec_z2 + (size_t)k * npar, z2, sizeof(double) * npar);
            propose(npar, old, z2, R, 1.0 / o->drscale, y2);
            if (out_of_bounds(npar, y2, low, upp)) {
                fl |= 16; counters[3]++;
            } else {
                double ss2 = orc_ss(c, N, t, ms2, pp7, y2); counters[0]++;
                double pri2 = prior_ss(npar, y2, pmu, psig);
                /* alpha(y2 -> y1), stage-1 form */
                double a32 = exp(-0.5 * ((ss1 - ss2) / sigma2 + pri1 - pri2));
                if (a32 > 1) a32 = 1;
                if (!(a32 >= 0)) a32 = 0;
                double l2 = -0.5 * ((ss2 - ss) / sigma2 + pri2 - pri);
                /* q1 = log q1(y1|y2)/q1(y1|x) with inv(R) of the stage-1 factor (mcmcstat qfun) */
                for (int i = 0; i < npar; ++i) scr[i] = y1[i] - y2[i];
                double n1 = norm2_dinvR(npar, scr, R, ybuf);
                for (int i = 0; i < npar; ++i) scr[i] = y1[i] - old[i];
                double n0 = norm2_dinvR(npar, scr, R, ybuf);
                double q1 = -0.5 * (n1 - n0);
                double a13 = exp(l2 + q1) * (1 - a32) / (1 - a12);
                if (a13 > 1) a13 = 1;
                int acc2;
                if (a13 >= 1) acc2 = 1;
                else {
                    double u = u2s ? u2s[k] : rng_unif(&rng);
                    if (rec_u2) rec_u2[k] = u;
                    acc2 = a13 > u;
                }
                if (acc2) { accept = 1; fl |= 2; newp = y2; ssn = ss2; prin = pri2; counters[2]++; }
            }
        }
        if (accept) {
            fl |= 1;
            memcpy(old, newp, sizeof(double) * npar); ss = ssn; pri = prin;
        } else { rej++; reju++; }
        memcpy(chain + (size_t)k * npar, old, sizeof(double) * npar);
        if (sschain) sschain[k] = ss;
        if (flags) flags[k] = fl;
        /* ---- sigma2 ~ inv-chi2(N0+N, (N0*S20+ss)/(N0+N)) */
        if (o->updatesigma) {
            double x2 = chi2s ? chi2s[k] : rng_chi2(&rng, o->N0 + o->Nobs);
            if (rec_chi2) rec_chi2[k] = x2;
            sigma2 = (o->N0 * o->S20 + ss) / x2;
        }
        s2chain[k] = sigma2;
        /* ---- adaptation */
        if (o->adaptint > 0 && isimu % o->adaptint == 0) {
            if (isimu < o->burnintime) {
                double rate = o->burnin_cumulative ? (double)rej / isimu : (double)reju / o->adaptint;
                double f = 1;
                if (rate > 0.95) f = 1 / o->burnin_scale;
                else if (rate < 0.05) f = o->burnin_scale;
                if (f != 1) for (size_t i = 0; i < np2; ++i) R[i] *= f;
                reju = 0;
            } else {
                /* covupd(chain(lasti+1:isimu,:),1,chaincov,chainmean,wsum) */
                if (!have_cov) {
                    int n = isimu - lasti;
                    for (int p = 0; p < npar; ++p) {
                        double s = 0;
                        for (int r = lasti; r < isimu; ++r) s += chain[(size_t)r * npar + p];
                        cmean[p] = s / n;
                    }
                    for (int p = 0; p < npar; ++p)
                        for (int q = 0; q <= p; ++q) {
                            double s = 0;
                            for (int r = lasti; r < isimu; ++r)
                                s += (chain[(size_t)r * npar + p] - cmean[p]) * (chain[(size_t)r * npar + q] - cmean[q]);
                            cov[(size_t)p * npar + q] = cov[(size_t)q * npar + p] = n > 1 ? s / (n - 1) : 0;
                        }
                    wsum = n; have_cov = 1;
                } else {
                    for (int r = lasti; r < isimu; ++r) {
                        const double *xi = chain + (size_t)r * npar;
                        double w = 1, wn = w + wsum;
                        for (int p = 0; p < npar; ++p) dd[p] = xi[p] - cmean[p];
                        double f1 = w / (wn - 1), f2 = wsum / wn;
                        for (int p = 0; p < npar; ++p)
                            for (int q = 0; q < npar; ++q) {
                                size_t ix = (size_t)p * npar + q;
                                cov[ix] = cov[ix] + f1 * (f2 * dd[p] * dd[q] - cov[ix]);
                            }
                        for (int p = 0; p < npar; ++p) cmean[p] += w / wn * dd[p];
                        wsum = wn;
                    }
                }
                lasti = isimu;
                int bad = 1;
                if (!o->qcovadj_always) bad = chol_upper(npar, cov, Rnew);          /* [Ra,is] = chol(upcov) */
                if (bad) {                                                          /* singular: "try to blow it" */
                    memcpy(tmpA, cov, sizeof(double) * np2);
                    for (int p = 0; p < npar; ++p) tmpA[(size_t)p * npar + p] += o->qcovadj;
                    bad = chol_upper(npar, tmpA, Rnew);
                }
                if (!bad) {
                    for (size_t i = 0; i < np2; ++i) R[i] = Rnew[i] * adascale;
                    counters[4]++;
                } else counters[5]++;
                reju = 0;
            }
        }
    }
    free(R); free(cov); free(tmpA); free(Rnew); free(cmean); free(old);
    return 0;
}

/* many chains, OpenMP over chains (the parfor analogue, TranscriptionCycleMCMC.m:161).
 * chain c uses cell chain_cell[c]; per-chain vectors are padded to ld = 7+Nmax.
 * Only summaries are returned: mean/popstd over rows [n_burn-1, nsimu) (MATLAB chain(n_burn:end,:)),
 * mean(s2chain) and std(sqrt(s2chain),1) over ALL rows (:302-303). */
void orc_run_chains(const orc_construct *c, const int *N, const long long *off, const double *t,
                    const double *ms2, const double *pp7, const orc_dram_opts *o, int n_burn,
                    int nchains, const int *chain_cell, const double *theta0, const double *qcov_diag,
                    const double *low, const double *upp, const double *pmu, const double *psig, int ld,
                    uint64_t seed, double *mean_out, double *std_out, double *sig_out,
                    long long *counters_out, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int ch = 0; ch < nchains; ++ch) {
        int cid = chain_cell[ch], n = N[cid], npar = 7 + n;
        orc_dram_opts oo = *o; oo.Nobs = 2.0 * n;
        double *chain = (double *)malloc(sizeof(double) * (size_t)o->nsimu * npar);
        double *s2 = (double *)malloc(sizeof(double) * o->nsimu);
        long long cnt[6];
        orc_dram(c, n, t + off[cid], ms2 + off[cid], pp7 + off[cid], &oo, theta0 + (size_t)ch * ld,
                 qcov_diag + (size_t)ch * ld, low + (size_t)ch * ld, upp + (size_t)ch * ld,
                 pmu + (size_t)ch * ld, psig + (size_t)ch * ld, 0, 0, 0, 0, 0,
                 seed + 0x9e3779b97f4a7c15ULL * (uint64_t)(ch + 1), 0, 0, 0, 0, 0, chain, s2, 0, 0, cnt);
        int r0 = n_burn - 1 < 0 ? 0 : n_burn - 1, nr = o->nsimu - r0;
        for (int p = 0; p < npar; ++p) {
            double s = 0;
            for (int r = r0; r < o->nsimu; ++r) s += chain[(size_t)r * npar + p];
            double mu = s / nr, q = 0;
            for (int r = r0; r < o->nsimu; ++r) { double e = chain[(size_t)r * npar + p] - mu; q += e * e; }
            mean_out[(size_t)ch * ld + p] = mu;
            std_out[(size_t)ch * ld + p] = sqrt(q / nr);
        }
        double s = 0, sr = 0;
        for (int r = 0; r < o->nsimu; ++r) { s += s2[r]; sr += sqrt(s2[r]); }
        double msr = sr / o->nsimu, q = 0;
        for (int r = 0; r < o->nsimu; ++r) { double e = sqrt(s2[r]) - msr; q += e * e; }
        sig_out[2 * ch] = sqrt(s / o->nsimu);
        sig_out[2 * ch + 1] = sqrt(q / o->nsimu);
        for (int i = 0; i < 6; ++i) counters_out[6 * (size_t)ch + i] = cnt[i];
        free(chain); free(s2);
    }
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
