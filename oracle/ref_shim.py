"""Import the UNMODIFIED reference (lmm077/SWEM at /root/reference) on CPU -- harness glue only.

/root/reference exists only in the authoring container, never on the GPU box, so nothing that
runs under ``-m gpu``, ``smoke()`` or ``bench.py`` may call this.  It is used by
``tests/golden/make_golden.py`` (to generate the committed fixtures) and by tests that are
skipped when the directory is absent.

Why a shim: ``methods/__init__.py`` pulls in tensorboardX (not installed),
``networks.py:8`` imports a name ``mod_resnet`` does not define, ``KeyEncoder`` torch.loads a
local ResNet checkpoint and the value encoders download weights.  We register empty package
modules so the ``__init__`` files are skipped, provide the missing name, and answer the weight
loads with random-init torchvision state dicts.  The reference sources are not touched or copied.
"""
from __future__ import annotations

import contextlib
import importlib
import importlib.util
import os
import sys
import types

REF = os.environ.get('SWEM_REFERENCE_DIR', '/root/reference')


def available() -> bool:
    return os.path.isfile(os.path.join(REF, 'methods', 'SWEM', 'modules.py'))


def load_modules():
    """Just ``methods/SWEM/modules.py`` (imports only math + torch) as a standalone module."""
    spec = importlib.util.spec_from_file_location('swem_reference_modules',
                                                  os.path.join(REF, 'methods', 'SWEM', 'modules.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@contextlib.contextmanager
def patched_weight_loads():
    """While active, the reference's weight loads (``torch.load`` of a local ResNet checkpoint in ``KeyEncoder``,
    ``model_zoo.load_url`` in the value encoders) are answered with random-init torchvision state dicts.  Everything is
    restored on exit, so later ``torch.load`` calls of the same process (the golden fixtures!) see the real function."""
    import torch
    import torchvision

    mr = importlib.import_module('methods.basic_modules.mod_resnet')
    real_load, real_url = torch.load, mr.model_zoo.load_url

    def fake_load(path, *a, **k):
        if path == '__r50__':
            return torchvision.models.resnet50(weights=None).state_dict()
        if path == '__r18__':
            return torchvision.models.resnet18(weights=None).state_dict()
        return real_load(path, *a, **k)

    torch.load = fake_load
    mr.model_zoo.load_url = lambda *a, **k: torchvision.models.resnet18(weights=None).state_dict()
    try:
        yield
    finally:
        torch.load = real_load
        mr.model_zoo.load_url = real_url


class _PatchedFactory:
    """Calls the reference class with the weight loads patched for the duration of the construction only."""

    def __init__(self, cls):
        self.cls = cls

    def __call__(self, *args, **kwargs):
        with patched_weight_loads():
            return self.cls(*args, **kwargs)


def load_swem():
    """Return (factory of the reference's SWEM, its modules module).  ``factory(cfg)`` builds the unmodified model with
    weight loading neutralised; nothing stays patched afterwards."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for name, rel in (('methods', 'methods'), ('methods.SWEM', 'methods/SWEM'),
                      ('methods.basic_modules', 'methods/basic_modules')):
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            pkg.__path__ = [os.path.join(REF, rel)]
            sys.modules[name] = pkg
    mr = importlib.import_module('methods.basic_modules.mod_resnet')
    mr.model_dirs = {'resnet50': '__r50__', 'resnet18': '__r18__'}     # networks.py:8 expects this name
    with patched_weight_loads():
        swem_mod = importlib.import_module('methods.SWEM.swem')
        modules_mod = importlib.import_module('methods.SWEM.modules')
    return _PatchedFactory(swem_mod.SWEM), modules_mod


def model_config(keydim=64, valdim=512, n_bases=128, n_iters=4, tau=0.05, topl=64,
                 single_obj=False, backbone='resnet50'):
    return types.SimpleNamespace(KEYDIM=keydim, VALDIM=valdim, NUM_BASES=n_bases, NUM_EM_ITERS=n_iters,
                                 EM_TAU=tau, TOPL=topl, SINGLE_OBJ=single_obj, BACKBONE=backbone)
