"""Builds libcoldbrew_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m gnn_tail_generalization_b200.build        # or: __graft_entry__.build()

nvcc cross-compiles without a GPU; the built library sits next to this file so that it travels with
the repo snapshot to the GPU box (it is git-ignored, not gpurun-ignored).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libcoldbrew_b200.so')
SOURCES = ['cb_graph.cu', 'cb_agg.cu', 'cb_elementwise.cu', 'cb_gemm.cu', 'cb_peer.cu', 'cb_topk.cu', 'cb_prep.cu']
NVCC_FLAGS = ['-std=c++17', '-O3', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC', '-shared']


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found (set NVCC or put /usr/local/cuda/bin on PATH)')


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, 'include', 'coldbrew_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library.  Returns the library path."""
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ['-I', os.path.join(ROOT, 'include'), '-I', CSRC, '-o', LIB + '.tmp'] \
        + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
        print(' '.join(cmd))
    subprocess.check_call(cmd)
    os.replace(LIB + '.tmp', LIB)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
