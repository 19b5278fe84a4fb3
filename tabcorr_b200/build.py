"""Build the CUDA extension in-tree: ``python -m tabcorr_b200.build``."""

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE = os.path.join(HERE, 'csrc', 'tabcorr_b200.cu')
OUTPUT = os.path.join(HERE, 'libtabcorr_b200.so')


def build(force=False, verbose=False):
    """Compile ``csrc/tabcorr_b200.cu`` (which includes ``csrc/*.cuh``) for sm_100a into
    ``libtabcorr_b200.so`` (skipped when the library is newer than its sources unless ``force``)."""
    header = os.path.join(os.path.dirname(HERE), 'include', 'tabcorr_b200.h')
    csrc = os.path.dirname(SOURCE)   # one translation unit: tabcorr_b200.cu includes the .cuh files
    sources = [header] + [os.path.join(csrc, name) for name in os.listdir(csrc)
                          if name.endswith(('.cu', '.cuh'))]
    if (not force and os.path.isfile(OUTPUT) and
            os.path.getmtime(OUTPUT) >= max(os.path.getmtime(path) for path in sources)):
        return OUTPUT
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    cmd = [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
           '-shared', '-Xcompiler', '-fPIC', '-o', OUTPUT, SOURCE]
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
        print(' '.join(cmd))
    subprocess.run(cmd, check=True)
    return OUTPUT


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose=True)
