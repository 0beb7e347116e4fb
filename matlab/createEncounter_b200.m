function traj = createEncounter_b200(self, sample_geo, tmax_s, varargin)
% CREATEENCOUNTER_B200  Drop-in body for @CorTerminalModel/createEncounter.m (lines 41-85: the four
% PropagateTrajectory chains, their concatenation and time sort) that runs on a B200 through emb_mex.
% sample_geo may be a struct ARRAY (a batch of encounters); the result is then n x 2.  The em-core smoothing of
% createEncounter.m:88-89 stays in MATLAB.  Reads only properties the classes have (CorTerminalModel.m:12-30,
% EncounterModel.m:5-70).  SOURCE ONLY (no MATLAB in the build image).
p = inputParser; addParameter(p, 'seed', 0); addParameter(p, 'first', 0); parse(p, varargin{:});
mdls = {self.mdlFwd1_1, self.mdlFwd1_2, self.mdlBck1_1, self.mdlBck1_2, self.mdlFwd2_1, self.mdlFwd2_2, self.mdlFwd2_3, ...
        self.mdlBck2_1, self.mdlBck2_2, self.mdlBck2_3};
hs = zeros(1, 10, 'uint64');
for i = 1:10
    m = mdls{i};
    % createEncounter.m:128-129: zero initial prior, stay prior on the transition tables (built by the reference's own
    % helpers; emb_handle adds them to the counts exactly as dbn_sample.m:124 does)
    hs(i) = emb_handle(m, 'dirichlet_initial', bn_dirichlet_prior(m.N_initial, 0), ...
                       'dirichlet_transition', setTransitionPriors(m.G_transition, m.r_transition, m.temporal_map, 1));
end
f = {'own_intent', 'own_distance', 'own_bearing', 'own_alt', 'own_heading', 'own_speed', ...
     'int_intent', 'int_distance', 'int_bearing', 'int_alt', 'int_heading', 'int_speed'};
n = numel(sample_geo);
geo = zeros(n, 12);
for j = 1:12, geo(:, j) = [sample_geo.(f{j})]'; end
lim = @(d) [d.minVel_ft_s; d.maxVel_ft_s; d.maxTurnRate_deg_s; d.maxAltitude_ft; d.maxVertRate_ft_s];
[t, len] = emb_mex('terminal_propagate', hs, p.Results.seed, p.Results.first, geo, tmax_s, [lim(self.dynLimits1), lim(self.dynLimits2)]);
tm = floor(tmax_s);
names = {'x_nm', 'y_nm', 'z_ft', 'heading_deg', 'v_ft_s'};
traj = repmat(struct(), n, 2);
for s = 1:n
    for ac = 1:2
        lo = tm + 1 - (double(len(s, 2 * ac)) - 1); hi = tm + 1 + double(len(s, 2 * ac - 1)) - 1;
        traj(s, ac).t_s = (lo:hi) - (tm + 1);
        for j = 1:5, traj(s, ac).(names{j}) = double(squeeze(t(s, lo:hi, ac, j)))'; end
    end
end
end
