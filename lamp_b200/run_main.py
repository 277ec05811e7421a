"""``python -m lamp_b200.run_main <reference_dir> [main.py arguments...]``

Runs the reference's unmodified ``main.py`` with the label-graph classes rebound to lamp_b200 (see compat.py).
Example (BASELINE cfg-1)::

    python -m lamp_b200.run_main /path/to/LaMP -dataset reuters -batch_size 32 -d_model 512 -d_inner_hid 512 \
        -n_layers_enc 2 -n_layers_dec 2 -n_head 4 -epoch 50 -dropout 0.2 -dec_dropout 0.2 -lr 0.0002 \
        -encoder graph -decoder graph -label_mask prior
"""
import os
import runpy
import sys


def main(argv):
    if len(argv) < 2:
        sys.exit(__doc__)
    ref = os.path.abspath(argv[1])
    from lamp_b200.compat import patch_reference
    if os.environ.get('LAMP_B200_NO_GPU_SHIM') == '1':
        # test aid for GPU-less boxes: the reference calls .cuda() unconditionally (train.py:34, test.py:37-47);
        # with this shim the run proceeds to the first lamp_b200 forward, which then refuses CPU tensors.
        import torch
        torch.Tensor.cuda = lambda self, *a, **k: self
    os.chdir(ref)
    patch_reference(ref)
    sys.argv = [os.path.join(ref, 'main.py')] + argv[2:]
    runpy.run_path(sys.argv[0], run_name='__main__')


if __name__ == '__main__':
    main(sys.argv)
