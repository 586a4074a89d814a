"""Build libs2vt_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libs2vt_b200.so')
SOURCES = ['s2vt_api.cu', 'ciderd.cu', 'rewards.cu', 'attention.cu', 'ingest.cpp', 'tfckpt.cpp']      # .cpp: host-only C ABI of include/s2vt_io.h
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC']


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return 'nvcc'


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.endswith('.o')] + \
        [os.path.join(HERE, '..', 'include', h) for h in ('s2vt.h', 's2vt_io.h')]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=True):
    """Compile every CUDA source of the package into one shared library.  Returns the library path."""
    if not force and not _stale():
        return LIB
    objs = []
    procs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(CSRC, os.path.splitext(src)[0] + '.o')
        if src.endswith('.cu'):
            cmd = [_nvcc()] + NVCC_FLAGS + os.environ.get('S2VT_NVCC_EXTRA', '').split() + ['-c', path, '-o', obj]     # e.g. -DS2VT_CHAIN_PROBE
        else:
            cmd = [os.environ.get('CXX', 'g++'), '-O3', '-std=c++17', '-fPIC', '-pthread', '-Wall', '-c', path, '-o', obj]
        if verbose:
            print(' '.join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out.decode())
            raise RuntimeError('compiler failed on %s' % src)
        elif verbose and out.strip():
            print(out.decode())
    cmd = [_nvcc(), '-shared', '-o', LIB] + objs
    if verbose:
        print(' '.join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv)
