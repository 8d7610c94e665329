function TranscriptionCycleMCMC(varargin)
%Drop-in for the reference's TranscriptionCycleMCMC.m with the per-cell DRAM
%fits running on the GPU engine (libtcmcmc, through tcmcmc_mex).
%
%Same variable arguments, same defaults and the same output files / struct
%layouts as the reference; 'numParPools' is the number of GPUs to use.
%The parfor body of the reference (per-cell set-up, mcmcrun, summaries,
%best-fit curves) is replaced by ONE call into the engine for all cells.
%New optional arguments: 'files' (cell array of dataset names: skips the
%listdlg dialog, for headless use), 'seed', 'saveChains'.

%% Variable inputs (defaults of the reference code)
fileDir = pwd; saveLoc = pwd; numParPools = 8; n_burn = 10000; n_steps = 20000;
ratePriorWidth = 50; t_start = 0; t_end = Inf; loadPrevious = false;
construct = 'P2P-MS2v5-LacZ-PP7v4'; files_arg = {}; seed = 20201028; saveChains = true;
prevFile = '';
for i = 1:length(varargin)
    if ~(ischar(varargin{i}) || isstring(varargin{i})), continue; end
    switch lower(char(varargin{i}))
        case 'filedir', fileDir = varargin{i+1};
        case 'saveloc', saveLoc = varargin{i+1};
        case 'numparpools', numParPools = varargin{i+1};
        case 'n_burn', n_burn = varargin{i+1};
        case 'n_steps', n_steps = varargin{i+1};
        case 'ratepriorwidth', ratePriorWidth = varargin{i+1};
        case 't_start', t_start = varargin{i+1};
        case 't_end', t_end = varargin{i+1};
        case 'loadprevious', loadPrevious = true;      %presence only, as in the reference
        case 'construct', construct = varargin{i+1};
        case 'files', files_arg = varargin{i+1};
        case 'seed', seed = varargin{i+1};
        case 'savechains', saveChains = varargin{i+1};
        case 'previousresults', prevFile = varargin{i+1};
    end
end
C = tcmcmc_constructs(construct);   %same table as GetFluorFromPolPos.m; errors on unknown names

%% Choose datasets (dialog only when a display exists and no 'files' given)
files = dir(fullfile(fileDir, '*.mat')); names = {files.name};
if ~isempty(files_arg)
    s = find(ismember(names, files_arg));
elseif usejava('desktop')
    [s, ~] = listdlg('PromptString', 'Select a dataset:', 'SelectionMode', 'multiple', 'ListString', names);
else
    s = 1:numel(names);
end

for k = 1:length(s)
    dat = load(fullfile(files(s(k)).folder, names{s(k)}));
    if ~isfield(dat, 'data'), continue; end
    data = dat.data; Ncells = length(data); DatasetName = data(1).name;
    prev = [];
    if loadPrevious
        P = load(prevFile); prev = P.MCMCresults;       %results file in this same layout
    end
    %per-cell set-up: truncation, x0, J0, bounds, priors (reference lines 163-255)
    cells = struct('time', {}, 'MS2', {}, 'PP7', {}); keep = [];
    for c = 1:Ncells
        t = data(c).time; i0 = find(t >= t_start, 1, 'first'); i1 = find(t < t_end, 1, 'last');
        if loadPrevious && ~any([prev.cell_index] == c), continue; end
        keep(end+1) = c; %#ok<AGROW>
        cells(end+1) = struct('time', t(i0:i1), 'MS2', data(c).MS2(i0:i1), 'PP7', data(c).PP7(i0:i1)); %#ok<AGROW>
    end
    n = numel(keep); Nmax = max(arrayfun(@(q) numel(q.time), cells)); ld = 7 + Nmax;
    x0 = zeros(ld, n); J0 = ones(ld, n); low = zeros(ld, n); upp = zeros(ld, n); mu = zeros(ld, n); sg = inf(ld, n);
    for q = 1:n
        t = cells(q).time; N = numel(t); np = 7 + N;
        if loadPrevious
            v0 = prev([prev.cell_index] == keep(q)).mean_v; v_step = 0.0000001; vl = v0 - 0.00001; vu = v0 + 0.00001;
        else
            v0 = 1 + 2*rand; v_step = 0.05; vl = 0; vu = 10;
        end
        x0(1:np, q) = [v0, 4*rand, 4*rand, 10, 5, rand, 15, normrnd(0, 3, 1, N)]';   %[v tau ton MS2b PP7b A R dR]
        J0(1:np, q) = [v_step, 0.1, t(end)-t(end-1), 1, 1, 0.05, 0.5, 0.5*ones(1, N)]';
        low(1:np, q) = [vl, 0, 0, 0, 0, 0, 0, -30*ones(1, N)]';
        upp(1:np, q) = [vu, 20, 10, 50, 50, 1, 40, 30*ones(1, N)]';
        sg(8:np, q) = ratePriorWidth;
    end
    opts = struct('n_steps', n_steps, 'n_burn', n_burn, 'numGPUs', numParPools, 'seed', seed, 'saveChains', saveChains);
    out = tcmcmc_mex('fit', C, cells, opts, x0, J0, low, upp, mu, sg);

    %% Pack the reference's structures (field order of the reference, lines 149-157)
    MCMCchain = struct('v_chain', {}, 'ton_chain', {}, 'A_chain', {}, 'tau_chain', {}, 'MS2_basal_chain', {}, ...
        'PP7_basal_chain', {}, 'R_chain', {}, 'dR_chain', {}, 's2chain', {});
    MCMCresults = struct('mean_v', {}, 'sigma_v', {}, 'mean_ton', {}, 'sigma_ton', {}, 'mean_A', {}, 'sigma_A', {}, ...
        'mean_tau', {}, 'sigma_tau', {}, 'mean_MS2_basal', {}, 'sigma_MS2_basal', {}, 'mean_PP7_basal', {}, ...
        'sigma_PP7_basal', {}, 'mean_R', {}, 'sigma_R', {}, 'mean_dR', {}, 'sigma_dR', {}, 'mean_sigma', {}, ...
        'sigma_sigma', {}, 'cell_index', {}, 'ApprovedFits', {});
    MCMCplot = struct('t_plot', {}, 'MS2_plot', {}, 'PP7_plot', {}, 'simMS2', {}, 'simPP7', {});
    for q = 1:n
        N = numel(cells(q).time); m = out.mean(:, q); sd = out.std(:, q);
        if saveChains
            ch = out.chain(:, :, q)';      %rows n_burn..n_steps
            MCMCchain(q) = struct('v_chain', ch(:,1), 'ton_chain', ch(:,3), 'A_chain', ch(:,6), 'tau_chain', ch(:,2), ...
                'MS2_basal_chain', ch(:,4), 'PP7_basal_chain', ch(:,5), 'R_chain', ch(:,7), 'dR_chain', ch(:,8:7+N), ...
                's2chain', out.s2chain(:, q));
        end
        R = struct('mean_v', m(1), 'sigma_v', sd(1), 'mean_ton', m(3), 'sigma_ton', sd(3), 'mean_A', m(6), 'sigma_A', sd(6), ...
            'mean_tau', m(2), 'sigma_tau', sd(2), 'mean_MS2_basal', m(4), 'sigma_MS2_basal', sd(4), 'mean_PP7_basal', m(5), ...
            'sigma_PP7_basal', sd(5), 'mean_R', m(7), 'sigma_R', sd(7), 'mean_dR', m(8:7+N)', 'sigma_dR', sd(8:7+N)', ...
            'mean_sigma', out.sig(1, q), 'sigma_sigma', out.sig(2, q), 'cell_index', keep(q), 'ApprovedFits', 0);
        if loadPrevious, R.ApprovedFits = prev([prev.cell_index] == keep(q)).ApprovedFits; end
        MCMCresults(q) = R;
        MCMCplot(q) = struct('t_plot', cells(q).time, 'MS2_plot', cells(q).MS2, 'PP7_plot', cells(q).PP7, ...
            'simMS2', out.simMS2(1:N, q)', 'simPP7', out.simPP7(1:N, q)');
    end
    filename = [date, '-', DatasetName];
    save(fullfile(saveLoc, [filename, '.mat']), 'MCMCresults', 'MCMCplot', 'DatasetName');
    if saveChains
        save(fullfile(saveLoc, [filename, '_RawChain.mat']), 'MCMCchain', '-v7.3');   %> 2 GiB safe
    end
end
disp(['MCMC analysis complete. Information stored in: ', saveLoc]);
end
