"""The C-ABI shared library loads and exports every symbol include/ptmcmc_b200.h declares; without a CUDA device the
engine refuses to start (there is no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re
import pytest
import torch
from ptmcmc_b200 import _capi as K
from ptmcmc_b200._lib import load, SO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ptmcmc_b200.h")


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ptg_[a-z_0-9]+)\s*\(", txt)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for s in ("ptg_create", "ptg_destroy", "ptg_set_space", "ptg_set_prior", "ptg_set_likelihood", "ptg_set_proposals",
              "ptg_inject_tapes", "ptg_inject_tape_marks", "ptg_init_from_prior", "ptg_init_states", "ptg_step", "ptg_step_host",
              "ptg_get_current", "ptg_get_history", "ptg_get_swap_stats", "ptg_get_trace", "ptg_checkpoint", "ptg_restore",
              "ptg_last_error", "ptg_abi_version"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(SO), "build the engine first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = load()
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert missing == []
    lib.ptg_abi_version.restype = C.c_int
    assert lib.ptg_abi_version() == K.ABI_VERSION


def test_config_struct_layout_matches_header():
    """ctypes mirror of ptg_config / ptg_proposal: field order and sizes as in the header"""
    txt = open(HEADER).read()
    body = re.search(r"typedef struct ptg_config \{(.*?)\} ptg_config;", txt, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"\b(?:int32_t|int64_t|uint64_t|double)\s+([a-zA-Z_0-9]+)\s*;", body)
    assert names == [f[0] for f in K.Config._fields_]
    assert C.sizeof(K.Config) == 12 * 4 + 5 * 8 + 8 + 8
    body = re.search(r"typedef struct ptg_proposal \{(.*?)\} ptg_proposal;", txt, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"([a-zA-Z_0-9]+)\s*[;,]", body)
    assert names == [f[0] for f in K.Proposal._fields_]


def test_bad_arguments_are_reported_not_crashed():
    lib = load()
    lib.ptg_create.restype = C.c_int
    lib.ptg_last_error.restype = C.c_char_p
    h = C.c_void_p()
    cfg = K.make_config(1, 4, 2)
    cfg.abi_version = 99
    assert lib.ptg_create(C.byref(cfg), C.byref(h)) == -1
    assert b"abi version" in lib.ptg_last_error()
    cfg = K.make_config(1, 4, 500)
    assert lib.ptg_create(C.byref(cfg), C.byref(h)) == -1
    assert lib.ptg_create(None, C.byref(h)) == -1
    lib.ptg_step.restype = C.c_int
    assert lib.ptg_step(None, C.c_int64(1)) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only box check")
def test_no_cpu_fallback_without_a_device():
    from ptmcmc_b200.engine import Engine
    with pytest.raises(K.CApiError, match="no CUDA device|CPU fallback"):
        Engine(K.make_config(1, 4, 2))
