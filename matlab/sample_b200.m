function [out_inits, out_events, out_samples, out_EME] = sample_b200(self, n_samples, sample_time, varargin)
% SAMPLE_B200  Drop-in body for UncorEncounterModel.sample (UncorEncounterModel.m:192-313) that
% runs the sampling on a B200 through emb_mex.  Same arguments ('seed', 'isQuantize500', 'layers'),
% same outputs.  SOURCE ONLY (no MATLAB in the build image).  A maintainer replaces the loop at
% UncorEncounterModel.m:244-307 by a call to this function, or puts this file on the path as an
% overloaded method.
p = inputParser;
addParameter(p, 'seed', nan); addParameter(p, 'isQuantize500', false); addParameter(p, 'layers', []);
parse(p, varargin{:});
persistent h fname
if isempty(h) || ~strcmp(fname, self.parameters_filename)
    h = emb_mex('load', self.parameters_filename, self.isOverwriteZeroBoundaries, self.idxZeroBoundaries);
    fname = self.parameters_filename;
end
info = emb_mex('info', h);
seed = p.Results.seed;
if isnan(seed), seed = randi(2^31 - 1); end          % 'seed' NaN: keep drawing from the caller's stream
find_lab = @(name) find(strcmp(self.labels_initial, name));
opts = struct('reject_mode', 1, 'idx_v', find_lab('"v"'), 'idx_dh', find_lab('"\dot h"'), 'idx_L', find_lab('"L"'), ...
              'is_quantize500', p.Results.isQuantize500, 'layers', p.Results.layers);
st = nan(1, self.n_initial);
for i = 1:self.n_initial, if ~isempty(self.start{i}), st(i) = self.start{i}; end, end
opts.start = st;
[out_inits, ~, values] = emb_mex('sample_tracks', h, seed, 0, n_samples, sample_time, opts);
% values: 4 x n x ceil(T/4) x n_timevarying  ->  out_samples{ii} n_initial x T  (events2samples.m:9-27)
tv = info.timevarying_vars;
values = reshape(permute(values, [2 4 1 3]), n_samples, numel(tv), []);   % n x tv x (4*nch)
out_samples = cell(n_samples, 1); out_events = cell(n_samples, 1); out_EME = cell(n_samples, 1);
map = info.temporal_map(:, 1)';
for ii = 1:n_samples
    d = repmat(out_inits(ii, :)', 1, sample_time);
    d(tv, :) = double(squeeze(values(ii, :, 1:sample_time)));
    out_samples{ii} = d;
    % events: every change of a time-varying value ([dt var value], dbn_hierarchical_sample.m:9-37).
    % Re-emissions that leave the value unchanged (zero bins) are not recoverable from the dense form;
    % the exact sparse list is the "next" row of DESIGN.md section 0.
    [var, col] = find(diff(d, 1, 2) ~= 0);
    [col, o] = sort(col); var = var(o);
    dt = diff([0; col]);
    out_events{ii} = [[dt, var, d(sub2ind(size(d), var, col + 1))]; sample_time - sum(dt), 0, 0];
    controls = events2controls(out_inits(ii, :), out_events{ii}, map);
    out_EME{ii} = controls;
end
end
