"""Build libodcuda.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'od_kernels.cu')
OUT = os.path.join(HERE, 'libodcuda.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-shared', '-Xcompiler', '-fPIC']
# (-split-compile was tried: the build is bound by the single-threaded front end, not by ptxas, and the split changes the
#  generated code -- the time-mode instantiation of step_spec_kernel came out 10 % slower on the B200)


def sources():
    d = os.path.join(HERE, 'csrc')
    return [os.path.join(d, f) for f in os.listdir(d)] + [os.path.join(HERE, '..', 'include', 'odcuda.h')]


def build(force=False, verbose=False):
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    if not os.path.exists(nvcc):
        nvcc = 'nvcc'
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(s) for s in sources()):
        return OUT
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', OUT, SRC]
    subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
