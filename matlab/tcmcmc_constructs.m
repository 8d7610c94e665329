function C = tcmcmc_constructs(construct)
%Construct table handed to the GPU engine: the same quantities, under the
%same names, as the if-block of the reference's GetFluorFromPolPos.m.
%Add your own reporter by adding a case (vectors = one entry per loop set).
switch construct
    case 'P2P-MS2v5-LacZ-PP7v4'
        C.L_MS2 = 6.626; C.L_PP7 = 6.626;          %kb, without the tau*v dwell term
        C.MS2_start = 0.024; C.MS2_end = 1.299; C.MS2_loopn = 24;
        C.PP7_start = 4.292; C.PP7_end = 5.758; C.PP7_loopn = 24;
    otherwise
        error('tcmcmc:construct', 'Unknown construct "%s": define it in tcmcmc_constructs.m', construct);
end
end
