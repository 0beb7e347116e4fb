function ok = verify_parity(varargin)
% VERIFY_PARITY  One-command check of the B200 sampler's oracle against the real, UNMODIFIED reference.
%
%     matlab -batch "addpath('<repo>/matlab'); verify_parity"
%     verify_parity('reference', '/path/to/em-model-manned-bayes', 'fixtures', '<repo>/tests/golden/matlab_parity')
%
% The CUDA sampler is tested bit for bit against a CPU restatement of the reference ("the oracle", oracle/ in the
% repository).  The reference ships no tests or golden vectors and MATLAB is not available where the sampler is built, so
% that restatement was never run against MATLAB itself.  This script does exactly that: it replays, through the shadow
% rand of matlab/inject/, the uniforms the oracle consumed (in the order in which the reference calls rand) into the
% reference's own classes and compares what the reference returns with what the oracle returned from the same uniforms
% (tests/golden/make_matlab_parity.py wrote both).  Cases:
%   uncor_fast    UncorEncounterModel.sample, uncor_1200code_v2p1  (fast branch of dbn_sample.m:95-166)
%   glider_slow   UncorEncounterModel.sample, glider_v1            (slow branch, dbn_sample.m:65-93)
%   terminal_geo  the loop of @CorTerminalModel/sample.m:29-77 on the terminal encounter geometry model (built from
%                 bn_sample / dediscretize because CorTerminalModel's constructor needs the unpublished trajectory models)
%   uncor_mt      mdl.sample(n, T, 'seed', 1) with MATLAB's own rng (no injection): pins the oracle's MT19937 provider
% Returns true when every case matches (out_inits and event values to 1e-12 relative, dt / var exactly) and prints one
% line per case.  Nothing in the reference is edited; matlab/inject is put ahead of it on the path for the tape cases only.
here = fileparts(mfilename('fullpath'));
p = inputParser;
addParameter(p, 'reference', getenv('AEM_DIR_BAYES'));
addParameter(p, 'fixtures', fullfile(here, '..', 'tests', 'golden', 'matlab_parity'));
parse(p, varargin{:});
ref = p.Results.reference;
fx = p.Results.fixtures;
assert(~isempty(ref) && isfolder(fullfile(ref, 'code', 'matlab')), ...
       'verify_parity: pass ''reference'', <checkout of em-model-manned-bayes> (or set AEM_DIR_BAYES)');
addpath(fullfile(ref, 'code', 'matlab'));
global EMB_TAPE EMB_TAPE_POS %#ok<GVMIS>
cases = {'uncor_fast', fullfile('model', 'uncor_1200code_v2p1.txt'); ...
         'glider_slow', fullfile('model', 'glider_v1.txt'); ...
         'terminal_geo', fullfile('model', 'correlated_terminal', 'terminalradar', 'terminal_v3_radar_encounter_model.txt'); ...
         'uncor_mt', fullfile('model', 'uncor_1200code_v2p1.txt')};
ok = true;
for c = 1:size(cases, 1)
    name = cases{c, 1};
    meta = load(fullfile(fx, [name '_meta.txt']));
    n = meta(1); T = meta(2); seed = meta(3);
    want_inits = load(fullfile(fx, [name '_inits.txt']));
    want_events = load_or_empty(fullfile(fx, [name '_events.txt']));
    tape = load(fullfile(fx, [name '_tape.txt']));
    file = fullfile(ref, cases{c, 2});
    use_tape = ~strcmp(name, 'uncor_mt');
    if use_tape
        addpath(fullfile(here, 'inject'));            % shadow rand ahead of the built-in
        EMB_TAPE = tape(:); EMB_TAPE_POS = 0;
    end
    try
        if strcmp(name, 'terminal_geo')
            got_inits = terminal_geometry(file, n);
            got_events = zeros(0, 4);
        else
            mdl = UncorEncounterModel('parameters_filename', file);
            if use_tape
                [got_inits, ev] = mdl.sample(n, T);                     % no 'seed': the tape is the stream
            else
                [got_inits, ev] = mdl.sample(n, T, 'seed', seed);       % MATLAB's own rng(seed,'twister')
            end
            got_events = zeros(0, 4);
            for ii = 1:n
                got_events = [got_events; ii * ones(size(ev{ii}, 1), 1), ev{ii}]; %#ok<AGROW>
            end
        end
        used = EMB_TAPE_POS;
    catch err
        cleanup_tape(here, use_tape);
        fprintf('%-13s ERROR  %s\n', name, err.message);
        ok = false;
        continue
    end
    cleanup_tape(here, use_tape);
    good = isequal(size(got_inits), size(want_inits)) && isequal(size(got_events), size(want_events));
    if good
        good = close_enough(got_inits, want_inits) && isequal(got_events(:, 1:3), want_events(:, 1:3)) && ...
               close_enough(got_events(:, 4), want_events(:, 4));
    end
    if good && use_tape
        good = used == numel(tape);                   % the reference consumed exactly the uniforms the oracle did
    end
    if good, verdict = 'PASS'; else, verdict = 'FAIL'; ok = false; end
    fprintf('%-13s %s   (%d samples, %d event rows, %d of %d uniforms consumed)\n', name, verdict, n, size(want_events, 1), ...
            used * use_tape, numel(tape) * use_tape);
end
if ok, disp('verify_parity: the oracle reproduces the reference on every case'); else, disp('verify_parity: MISMATCH'); end
end

function a = load_or_empty(f)
d = dir(f);
if isempty(d) || d.bytes == 0, a = zeros(0, 4); else, a = load(f); end
end

function cleanup_tape(here, use_tape)
global EMB_TAPE EMB_TAPE_POS %#ok<GVMIS>
EMB_TAPE = []; EMB_TAPE_POS = 0;
if use_tape, rmpath(fullfile(here, 'inject')); end
end

function tf = close_enough(a, b)
tf = all(abs(a(:) - b(:)) <= 1e-12 * max(abs(b(:)), 1e-300) | a(:) == b(:));
end

function outInits = terminal_geometry(file, nSamples)
% the loop of @CorTerminalModel/sample.m:29-77 with its default bounds (none) and the GENERIC speed limits of
% @CorTerminalModel/getDynamicLimits.m:15-17 (50 .. 506 ft/s), on an EncounterModel of the geometry file
mdl = EncounterModel('parameters_filename', file);
iOwn = find(strcmp(mdl.labels_initial, '"own_speed"'));
iInt = find(strcmp(mdl.labels_initial, '"int_speed"'));
outInits = zeros(nSamples, mdl.n_initial);
for ii = 1:nSamples
    isGood = false;
    while ~isGood
        initial = bn_sample(mdl.G_initial, mdl.r_initial, mdl.N_initial, mdl.dirichlet_initial, 1, mdl.start, mdl.order_initial);
        for kk = 1:numel(initial)
            if ~isempty(mdl.boundaries{kk})
                initial(kk) = dediscretize(initial(kk), mdl.boundaries{kk}, mdl.zero_bins{kk});
            end
        end
        isGood = initial(iOwn) <= 506 && initial(iOwn) >= 50 && initial(iInt) <= 506 && initial(iInt) >= 50;
    end
    outInits(ii, :) = initial;
end
end
