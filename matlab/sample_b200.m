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
[events, offsets, out_inits] = emb_mex('sample_events', h, seed, 0, n_samples, sample_time, opts);
% events: 4 x rows [dt; var; value; bin], rows offsets(ii)+1 : offsets(ii+1) belong to track ii -- the reference's own
% out_events{ii} (dbn_hierarchical_sample.m:9-37), built on the GPU in the reference's row order
out_samples = cell(n_samples, 1); out_events = cell(n_samples, 1); out_EME = cell(n_samples, 1);
map = info.temporal_map(:, 1)';
order = [find_lab('"\dot h"'), find_lab('"\dot \psi"'), find_lab('"\dot v"')];        % UncorEncounterModel.m:291-292
[~, cols] = ismember(order, map);
for ii = 1:n_samples
    ev = events(1:3, offsets(ii) + 1:offsets(ii + 1))';
    out_events{ii} = ev;
    out_samples{ii} = events2samples(out_inits(ii, :), ev);                               % events2samples.m:9-27
    controls = events2controls(out_inits(ii, :), ev, map);                                % events2controls.m:9-31
    controls = controls(:, [1, 1 + cols]);
    controls(:, 2) = controls(:, 2) / 60;                                                 % :295
    controls(:, 3) = deg2rad(controls(:, 3));                                             % :296
    controls(:, 4) = controls(:, 4) * 1.68780972222222;                                   % :297
    out_EME{ii} = controls;
end
end
