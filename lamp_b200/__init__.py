"""lamp_b200 -- B200-native implementation of LaMP's label-graph attention path.

The sub-modules mirror the reference package layout (``lamp.SubLayers`` ...), so the reference ``main.py`` can use
this package in place of ``lamp``.  Importing the package does not need a GPU; running any
forward does (there is no CPU path), and the native library ``liblamp_b200.so`` must have been built.
"""
from . import Constants  # noqa: F401
from . import utils  # noqa: F401
from . import ops  # noqa: F401
from . import SubLayers, Layers, Encoders, Decoders, Models  # noqa: F401
from .ops import set_default_precision  # noqa: F401
from .graphs import GraphedForward, GraphedTrainStep  # noqa: F401

__all__ = ['Constants', 'utils', 'ops', 'SubLayers', 'Layers', 'Encoders', 'Decoders', 'Models',
           'set_default_precision', 'GraphedForward', 'GraphedTrainStep']
