"""The C-ABI library loads and exports every symbol include/emb200.h declares (no GPU needed)."""
import ctypes as C
import os
import re

import pytest

from em_model_manned_bayes_b200 import _lib as L
from helpers import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "emb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(emb_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = L.lib()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(L.EXPORTED) == names
    assert lib.emb_abi_version() == 2


def test_struct_layouts_match_header_sizes():
    # sizes computed from the header's field list (natural alignment)
    assert C.sizeof(L.Rng) == 16
    assert C.sizeof(L.TrackOut) == 7 * 8
    assert C.sizeof(L.SampleOpts) == 24 * 4 + 4 * 4 + 2 * 24 * 8 + 4 + 4 + 8 * 2 * 8 + 4 * 3 + 4 + 8 + 8 + 8
    assert C.sizeof(L.ModelInfo) % 8 == 0


def test_rng_word_matches_oracle_stream():
    from oracle import philox as px
    lib = L.lib()
    for (seed, sample, attempt, purpose, index, sub, lane) in [(1, 0, 0, 1, 0, 0, 0), (2 ** 63 + 5, 12345678901, 3, 2, 77, 0, 3),
                                                             (42, 2 ** 40, 65535, 3, 599, 1, 2)]:
        got = lib.emb_rng_word(seed, sample, attempt, purpose, index, sub, lane)
        want = int(px.word(seed, sample, attempt, purpose, index, lane, sub=sub))
        assert got == want


def test_no_cpu_fallback_without_gpu(model_paths):
    """Sampling must fail loudly (EMB_E_CUDA) when no device is present -- never silently compute on the CPU."""
    lib = L.lib()
    if lib.emb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from em_model_manned_bayes_b200.model import EncounterModel
    m = EncounterModel(model_paths["balloon_v1"])
    with pytest.raises(L.EmbError) as ei:
        m.sample_initial(4, seed=1)
    assert ei.value.code == L.EMB_E_CUDA


def test_mex_gateway_source_matches_the_header():
    """matlab/emb_mex.cpp cannot be built here (no MATLAB), but it must at least compile against include/emb200.h:
    syntax- and type-check it with a declarations-only stand-in for mex.h (tests/stubs/mex.h)."""
    import subprocess
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "tests", "stubs"),
                        "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "matlab", "emb_mex.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src = open(os.path.join(ROOT, "matlab", "emb_mex.cpp")).read()
    for fn in ("emb_model_load", "emb_model_from_arrays", "emb_sample_track_events_packed", "emb_sample_tracks_multi", "emb_set_prior", "emb_sample_initial", "emb_sample_tracks", "emb_sample_track_events",
               "emb_terminal_propagate", "emb_terminal_screen", "emb_tracks_integrate"):
        assert fn + "(" in src, fn


def test_terminal_structs_match_header_sizes():
    assert C.sizeof(L.DynLimits) == 5 * 8
    assert C.sizeof(L.TerminalModels) == 10 * 8
    assert C.sizeof(L.TrajOut) == 2 * 8


MATLAB_DIR = os.path.join(ROOT, "matlab")
REF_MATLAB = "/root/reference/code/matlab"


def _classdef_members(path):
    """Property and method names declared by a MATLAB classdef file (and the method files of its @folder)."""
    import re
    names = set()
    block = None
    for ln in open(path, encoding="utf-8", errors="replace"):
        s = ln.strip()
        m = re.match(r"(properties|methods|events|enumeration)\b", s)
        if m:
            block = m.group(1)
            continue
        if s == "end" or s.startswith("end %"):
            continue
        if block == "properties":
            m = re.match(r"([A-Za-z]\w*)\s*(\(|=|;|$|\{|[A-Za-z])", s)
            if m and not s.startswith("%"):
                names.add(m.group(1))
        m = re.match(r"function\s+(?:\[?[^=]*\]?\s*=\s*)?(?:(?:get|set)\.)?([A-Za-z]\w*)", s)
        if m:
            names.add(m.group(1))
    folder = os.path.dirname(path)
    if os.path.basename(folder).startswith("@"):
        names |= {os.path.splitext(f)[0] for f in os.listdir(folder) if f.endswith(".m")}
    return names


@pytest.mark.skipif(not os.path.isdir(REF_MATLAB), reason="needs the reference checkout (build container only)")
def test_matlab_glue_only_reads_members_the_reference_classes_have():
    """Every `self.<name>` / `mdl.<name>` / `m.<name>` the shipped .m files read must be a property or method of the
    reference's classes (EncounterModel, UncorEncounterModel, CorTerminalModel) -- round 1's glue read
    parameters_filename / isOverwriteZeroBoundaries / idxZeroBoundaries, which the objects do not keep
    (@EncounterModel/EncounterModel.m:98-104 are constructor locals)."""
    import re
    members = set()
    for cls in ("EncounterModel", "UncorEncounterModel", "CorTerminalModel"):
        members |= _classdef_members(os.path.join(REF_MATLAB, "@" + cls, cls + ".m"))
    assert {"G_initial", "N_transition", "dirichlet_initial", "start", "temporal_map", "r_initial", "n_initial",
            "boundaries", "resample_rates", "mdlFwd1_1", "dynLimits1", "bounds_sample"} <= members
    assert "parameters_filename" not in members and "idxZeroBoundaries" not in members
    seen = 0
    for f in sorted(os.listdir(MATLAB_DIR)):
        if not f.endswith(".m"):
            continue
        src = "\n".join(ln.split("%")[0] for ln in open(os.path.join(MATLAB_DIR, f)))
        for obj, name in re.findall(r"\b(self|mdl|m)\.([A-Za-z]\w*)", src):
            assert name in members, "%s reads %s.%s, which the reference classes do not have" % (f, obj, name)
            seen += 1
    assert seen >= 30


@pytest.mark.skipif(not os.path.isdir(REF_MATLAB), reason="needs the reference checkout (build container only)")
def test_matlab_glue_calls_reference_functions_with_their_signatures():
    """The helper functions the glue calls must exist in the reference with the arity used."""
    import re
    for fn, nargs in (("events2samples", 2), ("events2controls", 3), ("bn_dirichlet_prior", 2), ("setTransitionPriors", 4)):
        head = open(os.path.join(REF_MATLAB, fn + ".m")).readline()
        m = re.search(r"%s\s*\(([^)]*)\)" % fn, head)
        assert m and len(m.group(1).split(",")) == nargs, head
    assert os.path.exists(os.path.join(REF_MATLAB, "@EncounterModelEvents", "EncounterModelEvents.m"))
