function run_reference(referenceDir, mcmcstatDir, outDir, numParPools, n_steps, seed)
% run_reference  Run the UNMODIFIED reference (GarciaLab/TranscriptionCycleInference) with mcmcstat on the path, time it,
% and dump what pins the sampler half of the oracle.  Needs MATLAB (Parallel Computing + Statistics toolboxes) or a recent
% GNU Octave with the statistics package; neither exists in the development image or on the GPU box, so this script could
% not be run there: the row "reference MATLAB path" of the baseline tables says "not measurable in this environment", and
% tests/test_mcmcstat_fixture.py activates when the fixture this script writes is dropped into tests/golden/.
%
%   run_reference('/path/to/TranscriptionCycleInference', '/path/to/mcmcstat', '/tmp/out', 8, 20000, 20201028)
%
% Part 1 (timing): TranscriptionCycleMCMC itself, exactly as shipped, on TestScripts/TestData.mat.  The reference picks its
%   datasets with a GUI dialog (listdlg) and shows a waitbar; baseline/shims/ shadows those two functions (select everything,
%   no window) — the reference's own files are not touched.  Prints seconds and chain-steps/s for numParPools workers.
% Part 2 (pin): the per-cell set-up of src/TranscriptionCycleMCMC.m:163-270 for cell 1, rng(seed), ONE mcmcrun call
%   (:273), saved with everything the loader test compares: x0, chain, s2chain, sschain, the results struct (R, qcov, cov,
%   mean, accepted, drscale, adascale, qcovadj, N0, S20, burnintime, burnscale ...), the mcmcstat version string, the RNG state.
if nargin < 4, numParPools = 8; end
if nargin < 5, n_steps = 20000; end
if nargin < 6, seed = 20201028; end
here = fileparts(mfilename('fullpath'));
addpath(fullfile(referenceDir, 'src'), fullfile(referenceDir, 'src', 'dependencies'), mcmcstatDir);
addpath(fullfile(here, 'shims'), '-begin');                      % headless listdlg / waitbar
if ~exist(outDir, 'dir'), mkdir(outDir); end

%% Part 1: the reference's own entry point, timed
dataDir = fullfile(outDir, 'data'); if ~exist(dataDir, 'dir'), mkdir(dataDir); end
copyfile(fullfile(referenceDir, 'TestScripts', 'TestData.mat'), dataDir);
n_burn = min(10000, floor(n_steps/2));
rng(seed);
t0 = tic;
TranscriptionCycleMCMC('fileDir', dataDir, 'saveLoc', outDir, 'numParPools', numParPools, 'n_burn', n_burn, 'n_steps', n_steps);
secs = toc(t0);
S = load(fullfile(dataDir, 'TestData.mat')); ncells = numel(S.data);
fprintf('reference: %d cells x %d steps on %d workers: %.1f s = %.0f chain-steps/s\n', ncells, n_steps, numParPools, secs, ncells*n_steps/secs);
timing = struct('seconds', secs, 'cells', ncells, 'n_steps', n_steps, 'n_burn', n_burn, 'numParPools', numParPools, ...
    'chain_steps_per_s', ncells*n_steps/secs, 'version', version); %#ok<NASGU>

%% Part 2: one mcmcrun call with a known seed (set-up copied from the reference's lines, values unchanged)
construct = 'P2P-MS2v5-LacZ-PP7v4'; ratePriorWidth = 50; cellNum = 1;
t = S.data(cellNum).time; MS2 = S.data(cellNum).MS2; PP7 = S.data(cellNum).PP7;      % t_start = 0, t_end = Inf: no truncation
data = struct; data.xdata = t; data.ydata = [MS2, PP7];                                % :179-181
ssfun = @(x, data) SumofSquaresFunction_TranscriptionCycleMCMC(construct, data, x);    % :186
rng(seed);
v0 = 1 + 2*rand; ton0 = 4*rand; A0 = rand; tau0 = 4*rand; MS2_basal0 = 10; PP7_basal0 = 5; R0 = 15;   % :200-206 (same draw order)
dR0 = normrnd(0, 3, 1, length(t));                                                      % :208
x0 = [v0, tau0, ton0, MS2_basal0, PP7_basal0, A0, R0, dR0];                           % :210
sigma2_0 = 1;                                                                           % :212
J0 = diag([0.05, 0.1, t(end)-t(end-1), 1, 1, 0.05, 0.5, 0.5*ones(size(dR0))]);        % :217-231
k = 0; names = {'v','tau','ton','MS2_basal','PP7_basal','A','R'};
lo = [0 0 0 0 0 0 0]; hi = [10 20 10 50 50 1 40];                                      % :242-249
params = cell(1, 7 + length(t));
for k = 1:7, params{k} = {names{k}, x0(k), lo(k), hi(k)}; end
for i = 1:length(dR0)                                                                   % :253-255
    params{7+i} = {['dR', num2str(i)], x0(7+i), -30, 30, 0, ratePriorWidth};
end
model = struct; model.ssfun = ssfun; model.sigma2 = sigma2_0; model.N = length(data.ydata);   % :257-260
options = struct; options.nsimu = n_steps; options.updatesigma = 1; options.qcov = J0;         % :263-270
options.burnintime = n_burn; options.adaptint = 100; options.method = 'dram'; options.verbosity = 0;
rngState = rng; %#ok<NASGU>
t1 = tic;
[results, chain, s2chain, sschain] = mcmcrun(model, data, params, options);            % :273
secs1 = toc(t1); %#ok<NASGU>
results.ssfun = []; results.modelfun = [];                                              % function handles do not survive save/load
mcmcstatVersion = '';
try, mcmcstatVersion = fileread(fullfile(mcmcstatDir, 'VERSION')); catch, end %#ok<NASGU>
save(fullfile(outDir, 'mcmcstat_run.mat'), 'results', 'chain', 's2chain', 'sschain', 'x0', 'J0', 'n_steps', 'n_burn', 'seed', ...
    'rngState', 'cellNum', 'construct', 'ratePriorWidth', 'timing', 'secs1', 'mcmcstatVersion', '-v7');
fprintf('wrote %s: copy it to tests/golden/mcmcstat_run.mat\n', fullfile(outDir, 'mcmcstat_run.mat'));
end
