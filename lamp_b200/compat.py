"""Drop the B200 path into the reference's own ``main.py`` without editing any reference file.

The reference has no plugin registry; its seam is Python class identity: ``main.py:7`` does
``from lamp.Models import LAMP`` and the layer classes import each other by name.  ``patch_reference()`` therefore
rebinds those names inside the reference's ``lamp`` package to the ``lamp_b200`` classes (same constructor /
``forward`` signatures and state-dict keys) and applies the two torch>=2 compatibility shims the reference needs
outside the label-graph path; ``python -m lamp_b200.run_main /path/to/reference <main.py args>`` then runs the
reference's byte-identical ``main.py``.
"""
import importlib
import os
import sys

import torch

_PATCHED = False


def _scoped_torch_load(reference_dir):
    """torch >= 2.6 defaults ``torch.load(weights_only=True)``; main.py:23 / :118 load pickled dicts.
    Only calls made FROM files of the reference tree get ``weights_only=False``; every other caller in the process keeps
    torch's safe default."""
    orig = torch.load
    root = os.path.realpath(reference_dir) + os.sep if reference_dir else None

    def load(*args, **kwargs):
        if 'weights_only' not in kwargs:
            caller = os.path.realpath(sys._getframe(1).f_code.co_filename)
            if root is None or caller.startswith(root):
                kwargs['weights_only'] = False
        return orig(*args, **kwargs)
    load.__wrapped__ = orig
    return load


def patch_reference(reference_dir=None):
    """Rebind ``lamp.{SubLayers,Layers,Encoders,Decoders,Models}`` classes to the lamp_b200 implementations."""
    global _PATCHED
    if reference_dir is not None and reference_dir not in sys.path:
        sys.path.insert(0, reference_dir)
    if not _PATCHED:
        torch.load = _scoped_torch_load(reference_dir)
        _PATCHED = True
    import lamp_b200
    ref = importlib.import_module('lamp')
    table = {
        'SubLayers': ['XavierLinear', 'ScaledDotProductAttention', 'MultiHeadAttention', 'PositionwiseFeedForward'],
        'Layers': ['EncoderLayer', 'DecoderLayer'],
        'Encoders': ['GraphEncoder'],
        'Decoders': ['GraphDecoder'],
        'Models': ['LAMP'],
    }
    for mod_name, names in table.items():
        ref_mod = importlib.import_module('lamp.' + mod_name)
        new_mod = getattr(lamp_b200, mod_name)
        for n in names:
            setattr(ref_mod, n, getattr(new_mod, n))
    # names imported *into* other reference modules at import time
    for holder, names in (('Layers', ['MultiHeadAttention', 'PositionwiseFeedForward']),
                          ('Encoders', ['EncoderLayer', 'DecoderLayer', 'ScaledDotProductAttention',
                                        'PositionwiseFeedForward', 'XavierLinear']),
                          ('Decoders', ['EncoderLayer', 'DecoderLayer', 'ScaledDotProductAttention',
                                        'PositionwiseFeedForward', 'XavierLinear']),
                          ('Models', ['EncoderLayer', 'DecoderLayer', 'ScaledDotProductAttention',
                                      'PositionwiseFeedForward', 'XavierLinear', 'GraphEncoder', 'GraphDecoder'])):
        ref_mod = importlib.import_module('lamp.' + holder)
        for n in names:
            if hasattr(ref_mod, n):
                for src in ('SubLayers', 'Layers', 'Encoders', 'Decoders'):
                    if hasattr(getattr(lamp_b200, src), n):
                        setattr(ref_mod, n, getattr(getattr(lamp_b200, src), n))
                        break
    # train.py:34 / test.py:47 build the multi-hot targets with a per-row Python loop on the host every step
    # (utils/utils.py:205-216); both look the function up on the module at call time, so it can be rebound as well
    try:
        ref_utils = importlib.import_module('utils.utils')
        ref_utils.get_gold_binary = lamp_b200.utils.get_gold_binary
    except ImportError:
        pass
    return ref
