"""Builds gsn_b200/libgsn_b200.so in-tree with nvcc for sm_100a.

    python -m gsn_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, 'csrc')
SO_PATH = os.path.join(_HERE, 'libgsn_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')

NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _deps():
    return sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(CSRC, '*.h')) + \
        [os.path.join(_HERE, '..', 'include', 'gsn_b200.h')]


def needs_build() -> bool:
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    return any(os.path.getmtime(f) > t for f in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return SO_PATH
    objs = []
    procs = []
    os.makedirs(os.path.join(_HERE, '_obj'), exist_ok=True)
    for src in sources():
        obj = os.path.join(_HERE, '_obj', os.path.basename(src)[:-3] + '.o')
        cmd = [NVCC] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f'--- nvcc {os.path.basename(src)} ---\n{out}\n')
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed')
    subprocess.check_call([NVCC, '-shared', '-o', SO_PATH] + objs + ['-lcudart'])
    return SO_PATH


if __name__ == '__main__':
    path = build(force='--force' in sys.argv, verbose='--verbose' in sys.argv)
    print(path)
