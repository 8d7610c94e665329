function h = waitbar(varargin)
% Headless stand-in for the progress window of src/TranscriptionCycleMCMC.m:139,143: no window.
h = [];
end
