"""Build libbcbf.so in-tree with nvcc for sm_100a (no torch dependency in the native code)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'lib', 'libbcbf.so')
SOURCES = ['api.cu', 'factor.cu', 'gram.cu', 'posterior.cu', 'ensemble.cu', 'socp.cu', 'ozaki.cu']
FLAGS = ['-shared', '-Xcompiler', '-fPIC', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3',
         '-std=c++17', '-diag-suppress', '177']


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, '..', 'include', 'bcbf.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + [os.path.join(CSRC, s) for s in SOURCES] + ['-o', OUT]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed building libbcbf.so')
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
