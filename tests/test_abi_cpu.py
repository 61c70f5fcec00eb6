"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol
include/crog_b200.h declares; the product path refuses to run without a B200 (no fallback)."""
import os
import re

import numpy as np
import pytest
import torch

from crog_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(L.SO_PATH):
        from crog_b200 import build

        build.build()
    return L.load()


def test_header_symbols_are_exported(built):
    hdr = open(os.path.join(ROOT, "include", "crog_b200.h")).read()
    declared = set(re.findall(r"\b(crog_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    for name in declared:
        assert hasattr(built, name), name
    assert built.crog_abi_version() == L.ABI_VERSION == 6
    assert built.crog_detect_workspace_bytes(4096, 416, 416, 5) > 0  # pure host arithmetic, no device needed


def test_gemm_struct_matches_header():
    hdr = open(os.path.join(ROOT, "include", "crog_b200.h")).read()
    body = hdr[hdr.index("typedef struct CrogGemm {"):hdr.index("} CrogGemm;")]
    fields = re.findall(r"(?:const\s+)?(?:void|float|int32_t|int64_t)\s*\*?\s*([a-zA-Z0-9_, ]+);", body)
    names = [n.strip() for f in fields for n in f.split(",")]
    assert names == [f[0] for f in L.CrogGemm._fields_]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from crog_b200 import synth
    from crog_b200.model import CROG
    from crog_b200.utils import grasp_eval as GE

    model = CROG(synth.default_cfg(17))
    with pytest.raises(L.CrogError):
        model(torch.zeros(1, 3, 416, 416), torch.zeros(1, 17, dtype=torch.long))
    with pytest.raises(L.CrogError):
        GE.detect_grasps(np.zeros((8, 8), np.float32), np.zeros((8, 8), np.float32), np.zeros((8, 8), np.float32),
                         np.zeros((8, 8), np.float32))
    with pytest.raises(L.CrogError):
        GE.calculate_iou([1, 2, 3, 20, 0], [1, 2, 3, 20, 0, 1])


def test_state_dict_contract_and_module_prefix():
    from crog_b200 import synth
    from crog_b200.model import build_crog

    cfg = synth.default_cfg(20)
    model, groups = build_crog(cfg)
    sd = synth.make_state_dict(cfg, 1, "init")
    assert set(model.state_dict()) == set(sd)
    assert all(tuple(model.state_dict()[k].shape) == tuple(v.shape) for k, v in sd.items())
    model.load_state_dict({"module." + k: v for k, v in sd.items()}, strict=True)
    assert torch.equal(model.state_dict()["proj.txt.weight"], sd["proj.txt.weight"])
    with pytest.raises(RuntimeError):
        bad = dict(sd); bad.pop("proj.txt.bias")
        model.load_state_dict(bad, strict=True)
    with pytest.raises(NotImplementedError):
        model.train()(torch.zeros(1, 3, 416, 416), torch.zeros(1, 20, dtype=torch.long))
