"""CPU oracle for the SDXL UNet training step (TEST INFRASTRUCTURE, NOT PRODUCT).

Everything under ``oracle/`` is a CPU restatement of the reference's algorithm for the
hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as
the checker / the CPU timing baseline.  The product package
(``aozora_sdxl_training_b200``) never imports it and has no CPU fallback.

Parity pinning status (see DESIGN.md "Oracle"):

* ``host_ref`` (ticket pool, loss-weight table, LR curve, seeded generators, Raven /
  Titan update, weighted MSE, exclusion) -- PINNED: checked against outputs of the
  reference's own code executed in the build container through ``oracle/ref_shim.py``;
  vectors frozen under ``tests/golden/`` by ``tests/golden/make_golden.py``.
* ``unet_ref`` / ``scheduler_ref`` (diffusers ``UNet2DConditionModel`` /
  ``DDPMScheduler``) -- PARITY UNPINNED: ``diffusers>=0.32.0`` (requirements.txt:2 of the
  reference, no lockfile) is third-party, not vendored and not installable here; the
  restatement follows its published SDXL architecture and is anchored only on the
  reference's own call sites (train.py:2613-2619, 2755-2761), the exact parameter
  census (1680 tensors / 2,567,463,684 parameters) and the round trip of every key
  through the reference's ``get_unet_key_mapping`` (train.py:2449-2465).
"""
