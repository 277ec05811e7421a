"""``python baseline/run_ref_main.py [main.py arguments...]`` -- the UNMODIFIED reference ``main.py`` (from
``baseline/_ref``) as plain PyTorch, on the GPU when one is visible, with nothing but the runtime shims of
SURVEY.md 8c (``Tensor.byte`` -> bool, ``torch.load(weights_only=False)``; ``Tensor.cuda`` -> identity only with
``-no_cuda``).  This is the "what does stock PyTorch give on the same box" arm next to
``python -m lamp_b200.run_main baseline/_ref ...`` (the same main.py with the label-graph classes rebound)."""
import functools
import os
import runpy
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from baseline import reference as ref  # noqa: E402


def main(argv):
    d = ref.ref_dir()
    if d is None:
        sys.exit('reference tree not available (baseline/_ref missing)')
    torch.Tensor.byte = lambda self, *a, **k: self.bool()
    torch.load = functools.partial(torch.load, weights_only=False)
    if '-no_cuda' in argv or not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    os.chdir(d)
    sys.path.insert(0, d)
    sys.argv = [os.path.join(d, 'main.py')] + argv
    runpy.run_path(sys.argv[0], run_name='__main__')


if __name__ == '__main__':
    main(sys.argv[1:])
