"""Import shim that makes the reference's own ``train.py`` importable (ORACLE ONLY).

Only usable where ``/root/reference`` exists (the build container); the GPU box never
calls this.  It fabricates the *names* ``train.py`` imports at load time from packages
that are not installed (``diffusers``, ``tomesd``: train.py:19-20, 36, 39-42); none of
them is on the arithmetic path of the functions we execute (ticket pool, loss table,
LR curve, generators, key map, Raven/Titan).  Used by ``tests/golden/make_golden.py``
to freeze known-answer vectors and by the container-only pinning tests.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("AOZORA_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "train.py"))


class _Inert:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return _Inert()

    def __call__(self, *a, **k):
        return _Inert()


def _stub(name, attrs=()):
    mod = types.ModuleType(name)
    for a in attrs:
        setattr(mod, a, type(a, (_Inert,), {}))
    sys.modules[name] = mod
    return mod


_train = None


def import_reference_train():
    """Return the reference's ``train`` module (real code, stubbed third-party imports)."""
    global _train
    if _train is not None:
        return _train
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    saved_argv = sys.argv
    sys.argv = sys.argv[:1]                      # TrainingConfig uses parse_known_args (train.py:277-279)
    try:
        if "diffusers" not in sys.modules:
            _stub("diffusers", ["StableDiffusionXLPipeline", "AutoencoderKL", "FlowMatchEulerDiscreteScheduler",
                                "DDPMScheduler", "UNet2DConditionModel"])
            _stub("diffusers.optimization").get_scheduler = lambda *a, **k: None
            _stub("diffusers.models")
            _stub("diffusers.models.attention_processor", ["AttnProcessor2_0", "FusedAttnProcessor2_0"])
        if "tomesd" not in sys.modules:
            _stub("tomesd")
        import train  # noqa: the reference's module
        _train = train
    finally:
        sys.argv = saved_argv
    return _train


def import_reference_optimizers():
    """The reference's RavenAdamW / TitanAdamW need no stubs (only torch)."""
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from training_utils.optimizers.raven import RavenAdamW
    from training_utils.optimizers.titan import TitanAdamW
    return RavenAdamW, TitanAdamW
