function MCMCchain = LoadRawChains(indexFile)
% LoadRawChains  Read the raw chains of a fit back into one 1 x Ncells struct array.
%
%   MCMCchain = LoadRawChains('<saveLoc>/<date>-<DatasetName>_RawChain.mat')
%
% Up to 2 GiB the engine writes the reference's single file (variable MCMCchain,
% TranscriptionCycleMCMC.m:377-378) and this function just loads it.  Beyond MAT v5's 2 GiB per
% variable the chains are split over <...>_RawChain_part<K>.mat (whole cells per part) and the file
% given here is the index: MCMCchainParts (file names), MCMCchainPartOfCell, nParts.
s = load(indexFile);
if isfield(s,'MCMCchain')
    MCMCchain = s.MCMCchain;
    return
end
folder = fileparts(indexFile);
MCMCchain = [];
for k = 1:s.nParts
    name = s.MCMCchainParts{k};
    p = load(fullfile(folder,name));
    if isempty(MCMCchain)
        MCMCchain = p.MCMCchain;
    else
        MCMCchain(p.firstCell:p.lastCell) = p.MCMCchain; %#ok<AGROW>
    end
end
end
