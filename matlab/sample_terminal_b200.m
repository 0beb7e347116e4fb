function [outInits, outSamples] = sample_terminal_b200(self, nSamples, varargin)
% SAMPLE_TERMINAL_B200  Drop-in body for @CorTerminalModel/sample.m:1-82 (encounter-geometry sampling with the bounds and
% speed rejection of :45-70) that runs on a B200 through emb_mex('sample_initial', ...) with reject_mode = 2.
% Same arguments ('seed'), same outputs: outInits nSamples x n_initial, outSamples cell of structs whose fields are the
% unquoted labels (:58-61).  Extra: 'starts', an nSamples x n_initial cell as returned by InitStartTerminal -- the whole set
% of start rows in ONE call instead of the loop of RUN_terminal.m:35-44.  SOURCE ONLY (no MATLAB in the build image).
p = inputParser;
addRequired(p, 'nSamples', @isnumeric);
addParameter(p, 'seed', nan, @isnumeric);
addParameter(p, 'starts', {}, @iscell);
parse(p, nSamples, varargin{:});
seed = p.Results.seed;
if ~isnan(seed) && ~isempty(seed)                                   % :19-22 (see sample_b200.m for the key convention)
    oldSeed = rng;
    rng(seed, 'twister');
    key = seed;
else
    key = randi(2^31 - 1);
end

h = emb_handle(self);
n = self.n_initial;
lo = -inf(1, n); hi = inf(1, n);
if ~isempty(self.bounds_sample)                                     % :45-53
    lo = self.bounds_sample(:, 1)'; hi = self.bounds_sample(:, 2)';
end
iOwn = find(strcmp(self.labels_initial, '"own_speed"'));
iInt = find(strcmp(self.labels_initial, '"int_speed"'));
lo(iOwn) = max(lo(iOwn), self.dynLimits1.minVel_ft_s); hi(iOwn) = min(hi(iOwn), self.dynLimits1.maxVel_ft_s);   % :64
lo(iInt) = max(lo(iInt), self.dynLimits2.minVel_ft_s); hi(iInt) = min(hi(iInt), self.dynLimits2.maxVel_ft_s);   % :65
st = nan(1, n);
for i = 1:n
    if ~isempty(self.start{i}), st(i) = self.start{i}; end
end
opts = struct('reject_mode', 2, 'box_lo', lo, 'box_hi', hi, 'start', st);
if ~isempty(p.Results.starts)
    rows = nan(nSamples, n);
    for i = 1:nSamples
        for k = 1:n
            if ~isempty(p.Results.starts{i, k}), rows(i, k) = p.Results.starts{i, k}; end
        end
    end
    opts.start_per_sample = rows;
end
[~, outInits] = emb_mex('sample_initial', h, key, 0, nSamples, opts);

outSamples = cell(nSamples, 1);
fieldNames = strrep(self.labels_initial, '"', '');                  % :59
for ii = 1:nSamples
    outSamples{ii} = cell2struct(num2cell(outInits(ii, :))', fieldNames(:), 1);
end
if ~isnan(seed) && ~isempty(seed)
    rng(oldSeed);                                                   % :79-81
end
end
