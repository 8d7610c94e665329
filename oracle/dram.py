"""NumPy restatement of mcmcstat::mcmcrun's DRAM loop with injectable randomness.
TEST INFRASTRUCTURE (see oracle/__init__.py).

mcmcstat (github.com/mjlaine/mcmcstat) is NOT vendored or version-pinned by the
reference (README.md:5); the only call is src/TranscriptionCycleMCMC.m:273 with the
configuration at :242-270.  This follows the published DRAM algorithm (Haario, Laine,
Mira & Saksman 2006, "DRAM: Efficient adaptive MCMC") with mcmcstat's defaults:
  * proposal  theta + randn(1,npar)*R, R = chol(qcov)      [fixture: jump var 0.4997 vs 0.5]
  * out-of-bounds => alpha = 0, ss = Inf, no ssfun call; DR is still attempted
  * one DR retry with R/5                                    [fixture: 5.007]
  * alpha13 = min(1, l(x,y2) q1(y2,y1)/q1(x,y1) (1-a(y2,y1))/(1-a(x,y1)))
  * sigma2 = (N0*S20+ss)/chi2(N0+N), N = 2*N_time, redrawn every step from the
    post-accept ss                                           [fixture: PIT uniform]
  * every adaptint steps: isimu < burnintime -> R scaled by /10 or *10 when the CUMULATIVE
    rejection rate is > 95 % / < 5 % (mcmcstat: `rejected > 0.95*isimu`); otherwise covupd over
    the rows since the last adaptation and R = chol(cov) * 2.4/sqrt(npar), with
    chol(cov + qcovadj*I) only when chol(cov) fails (mcmcstat: "try to blow it")  [unpinned;
    both are flags: burnin_cumulative, qcovadj_always]
PARITY UNPINNED beyond the bracketed fixture items.
"""
import numpy as np


def _prior_ss(th, mu, sig):
    return float(np.sum(((th - mu) / sig) ** 2))


def _d_invR_norm2(d, R):
    # || d * inv(R) ||^2 with R upper-triangular: solve y R = d
    y = np.linalg.solve(R.T, d)
    return float(y @ y)


def _chol_upper(A):
    """Upper Cholesky factor, or None when a pivot is not positive (MATLAB's second output of chol)."""
    n = A.shape[0]
    R = np.zeros_like(A)
    for j in range(n):
        for i in range(j + 1):
            s = A[i, j] - R[:i, i] @ R[:i, j]
            if i == j:
                if not s > 0:
                    return None
                R[j, j] = np.sqrt(s)
            else:
                R[i, j] = s / R[i, i]
    return R


class Recorded:
    """Randomness provider backed by recorded streams (row k = MCMC step k)."""

    def __init__(self, z1, u1, z2, u2, chi2):
        self.z1, self.u1, self.z2, self.u2, self.c2 = z1, u1, z2, u2, chi2

    def normal(self, k, stage, npar):
        return (self.z1 if stage == 1 else self.z2)[k]

    def uniform(self, k, stage):
        return (self.u1 if stage == 1 else self.u2)[k]

    def chi2(self, k, nu):
        return self.c2[k]


def make_streams(nsimu, npar, nu, seed):
    """Streams for replay tests (NumPy Generator; NOT the device's Philox)."""
    g = np.random.default_rng(seed)
    return dict(z1=g.standard_normal((nsimu, npar)), u1=g.random(nsimu),
                z2=g.standard_normal((nsimu, npar)), u2=g.random(nsimu),
                chi2=g.chisquare(nu, nsimu))


def dram(ssfun, theta0, qcov_diag, low, upp, pmu, psig, Nobs, nsimu, burnintime, rand,
         adaptint=100, drscale=5.0, ntry=2, adascale=None, qcovadj=1e-8, burnin_scale=10.0,
         N0=1.0, S20=1.0, sigma2_0=1.0, updatesigma=True, burnin_cumulative=True, qcovadj_always=False):
    npar = theta0.size
    if adascale is None:
        adascale = 2.4 / np.sqrt(npar)
    R = np.diag(np.sqrt(qcov_diag))
    chain = np.zeros((nsimu, npar)); s2chain = np.zeros(nsimu); sschain = np.zeros(nsimu)
    flags = np.zeros(nsimu, dtype=np.int32)
    old = theta0.copy(); ss = ssfun(old); pri = _prior_ss(old, pmu, psig); sigma2 = sigma2_0
    chain[0] = old; s2chain[0] = sigma2; sschain[0] = ss
    cov = None; cmean = None; wsum = 0.0; lasti = 0; rej = 0; reju = 0; nss = 1
    for k in range(1, nsimu):
        isimu = k + 1
        fl = 0; accept = False
        y1 = old + rand.normal(k, 1, npar) @ R
        if np.any(y1 < low) or np.any(y1 > upp):
            ss1, pri1, a12 = np.inf, 0.0, 0.0; fl |= 4
        else:
            ss1 = ssfun(y1); nss += 1; pri1 = _prior_ss(y1, pmu, psig)
            with np.errstate(over="ignore"):
                a12 = float(np.exp(-0.5 * ((ss1 - ss) / sigma2 + pri1 - pri)))
            if a12 <= 0: accept = False
            elif a12 >= 1: accept = True
            else: accept = a12 > rand.uniform(k, 1)
        newp, ssn, prin = y1, ss1, pri1
        if not accept and ntry >= 2:
            fl |= 8
            y2 = old + rand.normal(k, 2, npar) @ (R / drscale)
            if np.any(y2 < low) or np.any(y2 > upp):
                fl |= 16
            else:
                ss2 = ssfun(y2); nss += 1; pri2 = _prior_ss(y2, pmu, psig)
                with np.errstate(over="ignore", invalid="ignore"):
                    a32 = float(np.exp(-0.5 * ((ss1 - ss2) / sigma2 + pri1 - pri2)))
                a32 = min(1.0, a32) if a32 >= 0 else 0.0
                l2 = -0.5 * ((ss2 - ss) / sigma2 + pri2 - pri)
                q1 = -0.5 * (_d_invR_norm2(y1 - y2, R) - _d_invR_norm2(y1 - old, R))
                with np.errstate(over="ignore"):
                    a13 = min(1.0, float(np.exp(l2 + q1)) * (1 - a32) / (1 - a12))
                if a13 >= 1 or a13 > rand.uniform(k, 2):
                    accept = True; fl |= 2; newp, ssn, prin = y2, ss2, pri2
        if accept:
            fl |= 1; old = newp.copy(); ss = ssn; pri = prin
        else:
            rej += 1; reju += 1
        chain[k] = old; sschain[k] = ss; flags[k] = fl
        if updatesigma:
            sigma2 = (N0 * S20 + ss) / rand.chi2(k, N0 + Nobs)
        s2chain[k] = sigma2
        if adaptint > 0 and isimu % adaptint == 0:
            if isimu < burnintime:
                rate = rej / isimu if burnin_cumulative else reju / adaptint
                if rate > 0.95: R = R / burnin_scale
                elif rate < 0.05: R = R * burnin_scale
                reju = 0
            else:
                blk = chain[lasti:isimu]
                if cov is None:
                    n = blk.shape[0]; cmean = blk.mean(axis=0)
                    d = blk - cmean
                    cov = d.T @ d / (n - 1) if n > 1 else np.zeros((npar, npar)); wsum = float(n)
                else:
                    for xi in blk:           # covupd recursion, w = 1
                        wn = wsum + 1.0; d = xi - cmean
                        cov = cov + (1.0 / (wn - 1)) * ((wsum / wn) * np.outer(d, d) - cov)
                        cmean = cmean + d / wn; wsum = wn
                lasti = isimu
                Ra = None if qcovadj_always else _chol_upper(cov)       # [Ra,is] = chol(upcov)
                if Ra is None:                                          # singular: "try to blow it"
                    Ra = _chol_upper(cov + qcovadj * np.eye(npar))
                if Ra is not None:
                    R = Ra * adascale
                reju = 0
    return dict(chain=chain, s2chain=s2chain, sschain=sschain, flags=flags, nss=nss)
