"""In-tree build of the native library (``liblamp_b200.so``) for sm_100a.

``python -m lamp_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles without a GPU; the resulting
shared object sits next to this file so that it travels with the source tree and is found by ``_native.lib()``.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'lamp_capi.cu')
OUT = os.path.join(HERE, 'liblamp_b200.so')
DEPS = [os.path.join(HERE, 'csrc', f) for f in os.listdir(os.path.join(HERE, 'csrc'))] + \
       [os.path.join(os.path.dirname(HERE), 'include', 'lamp_b200.h')]

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared']


def nvcc_path() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found')


def up_to_date() -> bool:
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build(force: bool = False, verbose: bool = False, out: str = OUT, defines=()) -> str:
    """``defines`` + ``out``: debug variants (e.g. -DLAMP_ATTN_TRACE -> liblamp_b200_trace.so, scripts/attn_trace.py);
    the product library is always the default ``OUT`` without defines."""
    if not force and out == OUT and up_to_date():
        return OUT
    cmd = [nvcc_path()] + NVCC_FLAGS + [f'-D{d}' for d in defines] + (['-Xptxas', '-v'] if verbose else []) + \
        ['-o', out, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
    if verbose:
        print(r.stderr)
    return out


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
