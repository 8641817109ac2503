"""CPU tests: the C-ABI library loads without a GPU and exports every symbol include/aozora_b200.h declares;
the optimizer classes keep the reference's constructor contract; the product never imports the oracle."""
import ctypes
import os
import re

import pytest
import torch

from aozora_sdxl_training_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    decls = _lib.parse_header()
    assert len(decls) >= 35
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in decls:
        assert hasattr(lib, name), f"{name} declared in include/aozora_b200.h but not exported"
    assert _lib.load().aoz_abi_version() == 1
    assert _lib.query("aoz_mt_chunk_elems") == 16384


def test_argument_errors_are_reported_not_crashed():
    with pytest.raises(_lib.AozoraError) as e:
        _lib.call("aoz_add", 0, 0, 8, 0, 0)
    assert "null" in str(e.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "aozora_sdxl_training_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f


def test_optimizer_constructor_contract():
    from aozora_sdxl_training_b200.optimizers import RavenAdamW, TitanAdamW
    p = torch.nn.Parameter(torch.zeros(4))
    with pytest.raises(ValueError):
        RavenAdamW([p], lr=-1.0)
    with pytest.raises(ValueError):
        RavenAdamW([p], momentum_dtype=torch.float64)
    opt = RavenAdamW([{"params": [p], "lr_scale": 1.0}], lr=1e-4, betas=(0.9, 0.999), weight_decay=0.01, eps=1e-8,
                     debias_strength=0.3, momentum_dtype=torch.bfloat16)
    g = opt.param_groups[0]
    for key in ("lr", "lr_scale", "betas", "eps", "weight_decay", "debias_strength", "momentum_dtype"):
        assert key in g
    assert opt._momentum_dtype == torch.bfloat16 and opt.state_dict()["_momentum_dtype"] == torch.bfloat16
    assert opt.save_cpu_state() == {"_momentum_dtype": torch.bfloat16}
    # defaults of the reference classes (raven.py:28-33, titan.py:28-33)
    assert RavenAdamW([p]).defaults["betas"] == (0.9, 0.98) and RavenAdamW([p]).defaults["debias_strength"] == 0.9
    t = TitanAdamW([p])
    assert isinstance(t, TitanAdamW) and t.defaults["debias_strength"] == 1.0
    with pytest.raises(RuntimeError):
        TitanAdamW([p])                 # single-owner rule (titan.py:81-88)
    t.close()
    TitanAdamW([p]).close()
    # no CPU fallback: stepping a CPU parameter must fail loudly
    p.grad = torch.ones(4)
    with pytest.raises(_lib.AozoraError):
        opt.step()
