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


def test_build_optimizer_mirrors_create_optimizer():
    """train.py:2256-2270: package-default RAVEN_PARAMS (string momentum dtype, list betas) merged under the run's; missing keys
    fall back to eps 1e-8 / weight_decay 0.01 / debias_strength 1.0 only when absent from BOTH; lr = the LR curve's peak."""
    import torch
    from aozora_sdxl_training_b200 import train_loop
    from aozora_sdxl_training_b200.optimizers import RavenAdamW, TitanAdamW

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1 = torch.nn.Linear(4, 4)
            self.other = torch.nn.Linear(4, 4)

    class Cfg:
        LR_CUSTOM_CURVE = [[0.0, 0.0], [0.05, 8.0e-7], [1.0, 1.0e-7]]
        RAVEN_PARAMS = {"betas": [0.9, 0.999], "eps": 1e-8, "weight_decay": 0.01, "debias_strength": 0.3, "momentum_dtype": "bfloat16"}
        UNET_EXCLUDE_TARGETS = ["conv1"]

    net = Net()
    opt = train_loop.build_optimizer(Cfg, net)
    g = opt.param_groups[0]
    assert type(opt) is RavenAdamW and len(opt.param_groups) == 1 and g["lr_scale"] == 1.0
    assert g["lr"] == 8.0e-7 and g["betas"] == (0.9, 0.999) and g["eps"] == 1e-8 and g["weight_decay"] == 0.01
    assert g["debias_strength"] == 0.3 and g["momentum_dtype"] is torch.bfloat16 and opt._momentum_dtype is torch.bfloat16
    assert [p is q for p, q in zip(g["params"], net.other.parameters())] == [True, True]          # conv1 frozen by the keyword

    class Sparse:
        LEARNING_RATE = 2e-6
        RAVEN_PARAMS = {"momentum_dtype": "float32", "weight_decay": 0.0}

    g = train_loop.build_optimizer(Sparse, Net()).param_groups[0]
    assert g["lr"] == 2e-6 and g["momentum_dtype"] is torch.float32 and g["weight_decay"] == 0.0
    assert g["debias_strength"] == 0.3 and g["betas"] == (0.9, 0.999)                            # from the package defaults

    class Titan(Sparse):
        OPTIMIZER_TYPE = "Titan"
        TITAN_PARAMS = {"debias_strength": 1.0}

    topt = train_loop.build_optimizer(Titan, Net())
    assert isinstance(topt, TitanAdamW) and topt.param_groups[0]["debias_strength"] == 1.0
    topt.close()
