function [sel, ok] = listdlg(varargin)
% Headless stand-in for the dataset dialog of src/TranscriptionCycleMCMC.m:128-129: select every entry.
list = {};
for i = 1:2:numel(varargin)
    if strcmpi(varargin{i}, 'ListString'), list = varargin{i+1}; end
end
sel = 1:numel(list); ok = 1;
end
