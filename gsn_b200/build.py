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


PROFILE_SO_PATH = os.path.join(_HERE, 'libgsn_b200_prof.so')


def build(force: bool = False, verbose: bool = False, profile: bool = False) -> str:
    """profile=True builds libgsn_b200_prof.so with -DGSN_PROFILE_STAMPS (clock64 phase stamps inside the fused
    kernels, read by scripts/fm_debug.py); the shipped library carries no profiling hooks."""
    so_path = PROFILE_SO_PATH if profile else SO_PATH
    if not force and not profile and not needs_build():
        return SO_PATH
    objs = []
    procs = []
    obj_dir = os.path.join(_HERE, '_obj_prof' if profile else '_obj')
    os.makedirs(obj_dir, exist_ok=True)
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + '.o')
        cmd = [NVCC] + NVCC_FLAGS + (['-DGSN_PROFILE_STAMPS'] if profile else []) + \
            (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
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
    subprocess.check_call([NVCC, '-shared', '-o', so_path] + objs + ['-lcudart'])
    return so_path


if __name__ == '__main__':
    path = build(force='--force' in sys.argv, verbose='--verbose' in sys.argv, profile='--profile' in sys.argv)
    print(path)
