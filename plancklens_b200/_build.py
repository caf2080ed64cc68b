"""Builds plancklens_b200/csrc/libplk_b200.so with nvcc for sm_100a (in-tree, so it travels with the repo)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, 'csrc')
SO = os.path.join(CSRC, 'libplk_b200.so')
SOURCES = ['plk_api.cu']
HEADERS = ['plk_common.h', 'plk_tables.h', 'plk_legendre.cuh', 'plk_fft.cuh', 'plk_blas.cuh', 'plk_wigner.cuh', 'plk_rng.cuh']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared']


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(_HERE, '..', 'include', 'plk.h')]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', SO] + SOURCES
    subprocess.check_call(cmd, cwd=CSRC)
    return SO


if __name__ == '__main__':
    print(build(force=True, verbose=True))
