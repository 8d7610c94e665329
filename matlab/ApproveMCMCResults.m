function ApprovedFits = ApproveMCMCResults(varargin)
% ApproveMCMCResults  Headless, field-compatible curation of a results file of TranscriptionCycleMCMC.
%
% The reference's ApproveMCMCResults.m is an interactive figure loop that reads fields the current driver does not write
% (mean_dwell, dwell_chain, mean_R(2:end)) from hard-coded S:\ paths.  This version keeps its OUTPUT convention —
% MCMCresults(i).ApprovedFits = 1 approved / 0 uncurated / -1 rejected, written back into the same .mat file, which
% 'loadPrevious' carries into the next fit — on the fields the current driver DOES write, and needs no display:
%
%   ApproveMCMCResults('file', f, 'approve', [1 5 9], 'reject', 2, 'maxRhat', 1.1, 'maxSigma', 3, 'minESS', 200, ...
%                      'LoadPrevious', previousFile)
%
% 'approve' / 'reject': positions in MCMCresults.  Automatic rules for the fits still at 0: 'maxRhat' (MCMCdiagnostics.Rhat_max,
% needs 'numChains' > 1 at fit time), 'minESS' (smallest ESS of the seven head parameters), 'maxSigma' (mean_sigma).
% 'LoadPrevious': copy ApprovedFits from an earlier results file, matched on cell_index.  Same logic as
% transcriptioncycleinference_b200/curate.py.
file = ''; approve = []; reject = []; maxRhat = []; maxSigma = []; minESS = []; prevFile = '';
for i = 1:2:numel(varargin)
    switch lower(char(varargin{i}))
        case 'file', file = varargin{i+1};
        case 'approve', approve = varargin{i+1};
        case 'reject', reject = varargin{i+1};
        case 'maxrhat', maxRhat = varargin{i+1};
        case 'maxsigma', maxSigma = varargin{i+1};
        case 'miness', minESS = varargin{i+1};
        case 'loadprevious', prevFile = varargin{i+1};
    end
end
m = matfile(file, 'Writable', true);
MCMCresults = m.MCMCresults;
n = numel(MCMCresults);
app = [MCMCresults.ApprovedFits];
if ~isempty(prevFile)
    P = load(prevFile, 'MCMCresults'); pc = [P.MCMCresults.cell_index];
    for k = 1:n
        j = find(pc == MCMCresults(k).cell_index, 1);
        if ~isempty(j), app(k) = P.MCMCresults(j).ApprovedFits; end
    end
end
app(approve) = 1; app(reject) = -1;
auto = app == 0; verdict = zeros(1, n);
if ~isempty(maxRhat) || ~isempty(minESS)
    D = m.MCMCdiagnostics; dc = [D.cell_index];
    for k = 1:n
        j = find(dc == MCMCresults(k).cell_index, 1);
        if isempty(j), continue; end
        ok = true;
        if ~isempty(maxRhat), ok = ok && D(j).Rhat_max <= maxRhat; end
        if ~isempty(minESS) && ~isempty(D(j).ESS), ok = ok && min(D(j).ESS(1:7)) >= minESS; end
        verdict(k) = 2*ok - 1;
    end
end
if ~isempty(maxSigma)
    for k = 1:n
        if MCMCresults(k).mean_sigma > maxSigma, verdict(k) = -1; elseif verdict(k) == 0, verdict(k) = 1; end
    end
end
app(auto & verdict ~= 0) = verdict(auto & verdict ~= 0);
for k = 1:n, MCMCresults(k).ApprovedFits = app(k); end
m.MCMCresults = MCMCresults;            % as the reference does (ApproveMCMCResults.m:335)
ApprovedFits = app;
end
