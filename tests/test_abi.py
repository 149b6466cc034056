"""CPU: the C-ABI library loads, exports every symbol include/upb200.h declares, and agrees with params.py."""
import os
import re

import numpy as np

from drl_urban_planning_b200 import _lib, params as PL

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "upb200.h")).read()
    declared = set(re.findall(r"\b(upb_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"upb_ctx", "upb_config", "upb_status", "upb_clip_mode"}
    L = _lib.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in upb200.h but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    assert L.upb_abi_version() == 1


def test_param_layout_matches_python():
    slots = _lib.param_slots()
    assert len(slots) == len(PL.SLOTS) == 32
    for (name, off, rows, cols), s in zip(slots, PL.SLOTS.values()):
        assert name == s.name and off == s.offset
        shape = (rows, cols) if cols else (rows,)
        assert shape == s.shape, (name, shape, s.shape)
    assert _lib.lib().upb_num_params() == PL.NUM_PARAMS == _lib.UPB_NUM_PARAMS


def test_header_constants_match_python():
    hdr = open(os.path.join(ROOT, "include", "upb200.h")).read()
    consts = dict(re.findall(r"#define\s+(UPB_[A-Z_]+)\s+(\d+)", hdr))
    assert int(consts["UPB_NUM_PARAMS"]) == _lib.UPB_NUM_PARAMS
    assert int(consts["UPB_GRAD_STRIDE"]) == _lib.UPB_GRAD_STRIDE
    assert int(consts["UPB_STAT_OFFSET"]) == _lib.UPB_STAT_OFFSET
    assert int(consts["UPB_STAT_COUNT"]) == _lib.UPB_STAT_COUNT
    assert _lib.UPB_STAT_OFFSET + _lib.UPB_STAT_COUNT == _lib.UPB_GRAD_STRIDE


def test_errors_are_reported_not_thrown():
    import ctypes as C
    L = _lib.lib()
    n = C.c_uint64()
    rc = L.upb_pack_measure(1, None, 10, 10, 1, C.byref(n))
    assert rc != 0 and b"pack" in L.upb_last_error()


def test_cpulist_parser_of_the_numa_helper():
    from drl_urban_planning_b200.engine import _cpulist, bind_host_to_gpu_node
    assert _cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert _cpulist("") == []
    import torch
    if not torch.cuda.is_available():
        assert bind_host_to_gpu_node(0) is None      # no GPU topology to read: nothing is changed


def test_mlp_header_constants_match_python():
    hdr = open(os.path.join(ROOT, "include", "upb200.h")).read()
    consts = dict(re.findall(r"#define\s+(UPB_[A-Z_]+)\s+(\d+)", hdr))
    assert int(consts["UPB_MLP_NUM_PARAMS"]) == _lib.UPB_MLP_NUM_PARAMS == PL.MLP.num_params
    assert int(consts["UPB_MLP_GRAD_STRIDE"]) == _lib.UPB_MLP_GRAD_STRIDE
    assert int(consts["UPB_MLP_STAT_OFFSET"]) == _lib.UPB_MLP_STAT_OFFSET
    assert _lib.UPB_MLP_STAT_OFFSET + _lib.UPB_STAT_COUNT == _lib.UPB_MLP_GRAD_STRIDE
