"""Generate the golden fixtures by running the UNMODIFIED reference (``/root/reference``).

Run in the build container only (the GPU box has no ``/root/reference``)::

    python tests/golden/make_golden.py

The reference is imported as-is; the only adaptation is the three runtime shims of SURVEY.md
section 8c, applied by monkey-patching torch *before* the import (no reference file is edited):
``Tensor.cuda`` -> identity (CPU run), ``Tensor.byte`` -> ``Tensor.bool`` (torch>=2 ``masked_fill``),
``torch.load(weights_only=False)``.  Inputs/weights come from ``tests/golden/cases.py`` (seeded numpy),
so only the reference OUTPUTS are stored (``*.npz``, float32, big tensors row-subsampled).
"""
import functools
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
REF = os.environ.get('LAMP_REFERENCE', '/root/reference')

# ---- shims --------------------------------------------------------
torch.Tensor.cuda = lambda self, *a, **k: self
torch.Tensor.byte = lambda self, *a, **k: self.bool()
torch.load = functools.partial(torch.load, weights_only=False)
warnings.filterwarnings('ignore')
sys.path.insert(0, REF)
import lamp  # noqa: E402  (the reference package)
from lamp.SubLayers import MultiHeadAttention, ScaledDotProductAttention  # noqa: E402
from lamp.Models import LAMP  # noqa: E402

import cases  # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(1)  # deterministic reduction order for the stored fp32 values


def save(name, **arrays):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **{k: np.ascontiguousarray(v) for k, v in arrays.items()})
    print(f'{name}: ' + ', '.join(f'{k}{tuple(v.shape)}' for k, v in arrays.items()),
          f'-> {os.path.getsize(path) / 1024:.0f} KiB')


@torch.no_grad()
def gen_mha():
    for name, c in cases.MHA_CASES.items():
        p, q, kv, mask = cases.mha_inputs(c)
        d = c['D'] // c['H']
        m = MultiHeadAttention(c['H'], c['D'], d, d, dropout=0.1)
        m.load_state_dict(p, strict=True)
        m.eval()
        out, attn = m(q, kv, kv, attn_mask=None if mask is None else mask.contiguous())
        rs = c.get('row_stride', 1)
        arrays = dict(out=out[:, ::rs].numpy())
        if c.get('keep_attn', True):
            arrays['attn'] = attn[:, ::c.get('attn_row_stride', 1)].numpy()
        else:
            arrays['attn_rowsum'] = attn.sum(-1).numpy()
        save('mha_' + name, **arrays)


@torch.no_grad()
def gen_sdpa():
    rs = np.random.RandomState(5)
    n, lq, lk, d = 6, 33, 47, 32
    q = torch.from_numpy(rs.standard_normal((n, lq, d)).astype(np.float32))
    k = torch.from_numpy(rs.standard_normal((n, lk, d)).astype(np.float32))
    v = torch.from_numpy(rs.standard_normal((n, lk, d)).astype(np.float32))
    mask = torch.from_numpy(rs.rand(n, lq, lk) < 0.3)
    mask[:, :, 0] = False
    m = ScaledDotProductAttention(temperature=np.power(d, 0.5), dropout=0.1).eval()
    out, attn = m(q, k, v, attn_mask=mask)
    save('sdpa_small', out=out.numpy(), attn=attn.numpy())


@torch.no_grad()
def gen_models():
    for name, c in cases.MODEL_CASES.items():
        p, cfg, src_seq, src_pos, adj = cases.model_inputs(c)
        d = c['D'] // c['H']
        model = LAMP(c['V'] + 4, c['L'], c['T'], c['L'], n_layers_enc=c['n_enc'], n_layers_dec=c['n_dec'],
                     n_head=c['H'], n_head2=c['H'], d_word_vec=c['D'], d_model=c['D'], d_inner_hid=c['d_inner'],
                     d_k=d, d_v=d, dropout=0.2, dec_dropout=0.2, dec_dropout2=False, proj_share_weight=True,
                     encoder='graph', decoder='graph', enc_transform=c.get('enc_transform', ''),
                     no_enc_pos_embedding=not c.get('pos_enc', True),
                     label_adj_matrix=None if adj is None else adj.clone(),
                     label_mask=c['mask'])
        missing = model.load_state_dict(p, strict=True)
        model.eval()
        logits, enc_out, _ = model((src_seq, src_pos), None, None, None)
        logits2, _, enc_attns, dec_rest = model((src_seq, src_pos), None, None, None, return_attns=True)
        assert torch.equal(logits, logits2)
        dec_slf_attns, dec_enc_attns = dec_rest
        logits3, _, int_preds = model((src_seq, src_pos), None, None, None, int_preds=True)
        arrays = dict(logits=logits.numpy(), enc_output=enc_out[:, ::7].numpy(),
                      dec_slf_attn0=dec_slf_attns[0].numpy(),
                      dec_enc_attn_last=dec_enc_attns[-1][:, ::5].numpy(),
                      enc_slf_attn0=enc_attns[0][0][:, ::11, ::3].numpy(),
                      int_pred0=int_preds[0].numpy(), n_int_preds=np.array(len(int_preds)))
        save('model_' + name, **arrays)
        keys = sorted(model.state_dict().keys())
        with open(os.path.join(HERE, 'state_keys_' + name + '.txt'), 'w') as f:
            f.write('\n'.join(f'{k} {tuple(model.state_dict()[k].shape)}' for k in keys) + '\n')


if __name__ == '__main__':
    gen_sdpa()
    gen_mha()
    gen_models()
