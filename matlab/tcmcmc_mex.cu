// tcmcmc_mex.cu — MEX gateway: MATLAB <-> libtcmcmc C ABI (include/tcmcmc.h).
//
// Build (on a machine with MATLAB + CUDA; NOT buildable in the development container, which has no
// mex.h — everything testable lives behind the C ABI, this file only marshals):
//     mexcuda -I../include tcmcmc_mex.cu -L../transcriptioncycleinference_b200 -ltcmcmc
//
// MATLAB call (from matlab/TranscriptionCycleMCMC.m, replacing the parfor body of the reference,
// src/TranscriptionCycleMCMC.m:161-357):
//
//   out = tcmcmc_mex('fit', construct, cells, opts, x0, J0diag, low, upp, prior_mu, prior_sig)
//
//   construct  struct: L_MS2, L_PP7, MS2_start, MS2_end, MS2_loopn, PP7_start, PP7_end, PP7_loopn
//              (the quantities of src/GetFluorFromPolPos.m:18-30; vectors = one entry per loop set)
//   cells      1 x Ncells struct array with fields time, MS2, PP7 (already truncated to [t_start,t_end))
//   opts       struct: n_steps, n_burn, numGPUs, seed, saveChains  (+ optional mcmcstat overrides)
//   x0 .. prior_sig   npar_max x Ncells double (column c = chain of cell c, zero padded)
//
//   out        struct: mean, std (npar_max x Ncells), sig (2 x Ncells), counters (16 x Ncells, int64),
//              chain ((n_steps-n_burn+1) x npar_max x Ncells), s2chain (n_steps x Ncells)
//              [chain/s2chain only when opts.saveChains], simMS2/simPP7 (Nmax x Ncells)
#include <cstring>
#include <vector>

#include "mex.h"
#include "tcmcmc.h"

static double field_scalar(const mxArray *s, const char *name, double dflt)
{
    const mxArray *f = mxGetField(s, 0, name);
    return f ? mxGetScalar(f) : dflt;
}

static void fill_vec(const mxArray *s, const char *name, double *dst, int &n)
{
    const mxArray *f = mxGetField(s, 0, name);
    if (!f) mexErrMsgIdAndTxt("tcmcmc:construct", "construct.%s is missing", name);
    n = (int)mxGetNumberOfElements(f);
    if (n < 1 || n > TC_MAX_SETS) mexErrMsgIdAndTxt("tcmcmc:construct", "construct.%s: 1..%d loop sets", name, TC_MAX_SETS);
    std::memcpy(dst, mxGetPr(f), sizeof(double) * n);
}

static void check(int rc)
{
    if (rc < 0) mexErrMsgIdAndTxt("tcmcmc:engine", "%s", tc_last_error());
}

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    if (nrhs != 10 || !mxIsChar(prhs[0])) mexErrMsgIdAndTxt("tcmcmc:usage", "tcmcmc_mex('fit', construct, cells, opts, x0, J0, low, upp, mu, sig)");
    // ---- construct table
    tc_construct c;
    std::memset(&c, 0, sizeof(c));
    c.L_ms2 = field_scalar(prhs[1], "L_MS2", 0);
    c.L_pp7 = field_scalar(prhs[1], "L_PP7", 0);
    int n = 0, n2 = 0;
    fill_vec(prhs[1], "MS2_start", c.ms2_start, n);
    fill_vec(prhs[1], "MS2_end", c.ms2_end, n2);
    fill_vec(prhs[1], "MS2_loopn", c.ms2_loopn, n2);
    fill_vec(prhs[1], "PP7_start", c.pp7_start, n2);
    fill_vec(prhs[1], "PP7_end", c.pp7_end, n2);
    fill_vec(prhs[1], "PP7_loopn", c.pp7_loopn, n2);
    c.nsets = n;
    // ---- cells -> packed arrays
    const int ncells = (int)mxGetNumberOfElements(prhs[2]);
    std::vector<int32_t> N(ncells);
    std::vector<int64_t> off(ncells + 1, 0);
    std::vector<double> t, ms2, pp7;
    int Nmax = 0;
    for (int i = 0; i < ncells; ++i) {
        const mxArray *ft = mxGetField(prhs[2], i, "time"), *f1 = mxGetField(prhs[2], i, "MS2"), *f2 = mxGetField(prhs[2], i, "PP7");
        N[i] = (int32_t)mxGetNumberOfElements(ft);
        off[i + 1] = off[i] + N[i];
        Nmax = N[i] > Nmax ? N[i] : Nmax;
        t.insert(t.end(), mxGetPr(ft), mxGetPr(ft) + N[i]);
        ms2.insert(ms2.end(), mxGetPr(f1), mxGetPr(f1) + N[i]);
        pp7.insert(pp7.end(), mxGetPr(f2), mxGetPr(f2) + N[i]);
    }
    // ---- options: mcmcstat defaults + the reference's configuration, then the caller's values
    tc_mcmc_opts o;
    tc_opts_default(&o);
    o.nsimu = (int)field_scalar(prhs[3], "n_steps", o.nsimu);
    o.burnintime = o.n_burn = (int)field_scalar(prhs[3], "n_burn", o.n_burn);
    o.ngpus = (int)field_scalar(prhs[3], "numGPUs", 1);            // 'numParPools' reinterpreted
    o.seed = (uint64_t)field_scalar(prhs[3], "seed", (double)o.seed);
    o.store_chain = (int)field_scalar(prhs[3], "saveChains", 1);
    o.adaptint = (int)field_scalar(prhs[3], "adaptint", o.adaptint);
    o.drscale = field_scalar(prhs[3], "drscale", o.drscale);
    o.qcovadj = field_scalar(prhs[3], "qcovadj", o.qcovadj);
    o.N0 = field_scalar(prhs[3], "N0", o.N0);
    o.layout = (int32_t)field_scalar(prhs[3], "layout", o.layout);   // TC_LAYOUT_AUTO; 1 forces the large-series layout
    int ndev = 0;
    check(tc_device_count(&ndev));
    if (o.ngpus > ndev) o.ngpus = ndev;
    const int ld = (int)mxGetM(prhs[4]);                             // npar_max rows, one column per cell
    if ((int)mxGetN(prhs[4]) != ncells || ld < 7 + Nmax) mexErrMsgIdAndTxt("tcmcmc:dims", "x0 must be (7+max N) x Ncells");
    // ---- run
    tc_cells *cells = nullptr;
    std::vector<int32_t> devs(o.ngpus);
    for (int d = 0; d < o.ngpus; ++d) devs[d] = d;
    check(tc_cells_create(&c, ncells, N.data(), off.data(), t.data(), ms2.data(), pp7.data(), o.ngpus, devs.data(), &cells));
    std::vector<int32_t> chain_cell(ncells);
    for (int i = 0; i < ncells; ++i) chain_cell[i] = i;
    const char *names[] = {"mean", "std", "sig", "counters", "chain", "s2chain", "simMS2", "simPP7"};
    plhs[0] = mxCreateStructMatrix(1, 1, 8, names);
    mxArray *mean = mxCreateDoubleMatrix(ld, ncells, mxREAL), *sd = mxCreateDoubleMatrix(ld, ncells, mxREAL);
    mxArray *sig = mxCreateDoubleMatrix(2, ncells, mxREAL);
    mxArray *cnt = mxCreateNumericMatrix(TC_NCOUNTERS, ncells, mxINT64_CLASS, mxREAL);
    mxArray *chain = nullptr, *s2 = nullptr;
    if (o.store_chain) {
        // MATLAB is column-major: [ld x nrows x ncells] here is the engine's row-major [ncells][nrows][ld]
        const mwSize dims[3] = {(mwSize)ld, (mwSize)(o.nsimu - o.n_burn + 1), (mwSize)ncells};
        chain = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);
        s2 = mxCreateDoubleMatrix(o.nsimu, ncells, mxREAL);
    }
    const int rc = tc_mcmc_run(cells, &o, ncells, chain_cell.data(), nullptr, ld, mxGetPr(prhs[4]), mxGetPr(prhs[5]),
                               mxGetPr(prhs[6]), mxGetPr(prhs[7]), mxGetPr(prhs[8]), mxGetPr(prhs[9]), mxGetPr(mean),
                               mxGetPr(sd), mxGetPr(sig), (int64_t *)mxGetData(cnt), chain ? mxGetPr(chain) : nullptr,
                               s2 ? mxGetPr(s2) : nullptr, nullptr);
    // best-fit curves at the posterior means on the RAW grid (src/TranscriptionCycleMCMC.m:307-309)
    mxArray *sim1 = mxCreateDoubleMatrix(Nmax, ncells, mxREAL), *sim2 = mxCreateDoubleMatrix(Nmax, ncells, mxREAL);
    int rc2 = rc < 0 ? rc : tc_forward(cells, ncells, chain_cell.data(), mxGetPr(mean), ld, 1, mxGetPr(sim1), mxGetPr(sim2), Nmax);
    tc_cells_destroy(cells);
    check(rc2);
    mxSetField(plhs[0], 0, "mean", mean); mxSetField(plhs[0], 0, "std", sd); mxSetField(plhs[0], 0, "sig", sig);
    mxSetField(plhs[0], 0, "counters", cnt);
    if (chain) { mxSetField(plhs[0], 0, "chain", chain); mxSetField(plhs[0], 0, "s2chain", s2); }
    mxSetField(plhs[0], 0, "simMS2", sim1); mxSetField(plhs[0], 0, "simPP7", sim2);
}
