function h = emb_handle(mdl, varargin)
% EMB_HANDLE  libemb200 model handle for an EncounterModel object (or a struct with the same fields), built from the
% properties the class really has (@EncounterModel/EncounterModel.m:5-70) -- the object does not keep its file name, and it
% can be constructed from arrays without any file (:106-114).
%
%   h = emb_handle(mdl)                                   priors: mdl.dirichlet_initial / mdl.dirichlet_transition
%   h = emb_handle(mdl, 'dirichlet_transition', alpha)    override (createEncounter.m:128-129 builds its own tables)
%
% The tables handed to the library are the weights select_random.m:17 sums, N{i} + alpha{i} (bn_sample.m:54,
% dbn_sample.m:75,124), so any prior -- constant, 'dbe', the stay prior of setTransitionPriors.m or a hand-made
% dirichlet cell -- is reproduced exactly; the library's own prior stays 0.
% Handles are cached on a fingerprint of the tables (the properties are publicly settable, so the object identity is not
% enough); emb_handle('clear') frees them.  SOURCE ONLY (no MATLAB in the build image).
persistent cache
if isempty(cache), cache = containers.Map('KeyType', 'char', 'ValueType', 'uint64'); end
if ischar(mdl) && strcmp(mdl, 'clear')
    ks = keys(cache);
    for i = 1:numel(ks), emb_mex('free', cache(ks{i})); end
    cache = containers.Map('KeyType', 'char', 'ValueType', 'uint64');
    h = uint64(0);
    return
end
p = inputParser;
addParameter(p, 'dirichlet_initial', mdl.dirichlet_initial);
addParameter(p, 'dirichlet_transition', mdl.dirichlet_transition);
parse(p, varargin{:});
w_initial = add_prior(mdl.N_initial, p.Results.dirichlet_initial);
w_transition = add_prior(mdl.N_transition, p.Results.dirichlet_transition);
flat = @(c) cell2mat(cellfun(@(x) x(:), c(:), 'UniformOutput', false));
wi = flat(w_initial); wt = flat(w_transition);
rates = mdl.resample_rates(:);
key = sprintf('%d|%d|%.17g|%.17g|%.17g|%.17g|%s', numel(wi), numel(wt), sum(wi), sum(wt), sum(wi .* (1:numel(wi))'), ...
              sum(wt .* (1:numel(wt))'), mat2str([rates' cellfun(@numel, mdl.boundaries)]));
if isKey(cache, key), h = cache(key); return, end
h = emb_mex('from_arrays', logical(mdl.G_initial), double(mdl.r_initial(:)), wi, logical(mdl.G_transition), ...
            double(mdl.r_transition(:)), wt, double(mdl.temporal_map), mdl.boundaries, rates);
cache(key) = h;
end

function w = add_prior(N, alpha)
w = N;
for i = 1:numel(N)
    if ~isempty(N{i}) && numel(alpha) >= i && ~isempty(alpha{i}), w{i} = N{i} + alpha{i}; end
end
end
