% bench_matlab.m -- times the reference's own CPU path on BASELINE.json configs[0]
% (UncorEncounterModel on model/uncor_1200code_v2p1.txt, 1,000 tracks x 300 s, seed 1).
% Not runnable in the build image (no MATLAB); shipped so that anyone with MATLAB >= R2020a can put
% the reference's number next to bench.py's.  Run from the reference checkout after startup_bayes.
mdl = UncorEncounterModel('parameters_filename', [getenv('AEM_DIR_BAYES') filesep 'model' filesep 'uncor_1200code_v2p1.txt']);
n = 1000; T = 300;
mdl.sample(10, T, 'seed', 1);                 % warm-up (JIT, file cache)
t0 = tic; mdl.sample(n, T, 'seed', 1); s = toc(t0);
fprintf('{"impl": "matlab", "metric": "sampled track-timesteps/sec", "value": %.6g, "unit": "track-timesteps/s", "cores": 1, "sample": "%d tracks x %d s"}\n', n * T / s, n, T);
