"""Build recipe for ``liborbit_b200.so`` (nvcc, sm_100a only).

The library is built IN-TREE next to this file so that it travels with the
repository snapshot to the GPU box; it is git-ignored.  ``python -m
nbody_orbit_analysis_b200._build`` (or ``__graft_entry__.build()``) rebuilds it
when a source is newer than the library.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'liborbit_b200.so')
SOURCES = ['oa_api.cu', 'oa_track.cu', 'oa_select.cu', 'oa_bulk.cu',
           'oa_sort.cu', 'oa_synth.cu', 'oa_segment.cu', 'oa_join.cu',
           'oa_pjoin.cu']
HEADERS = [os.path.join(CSRC, 'oa_common.cuh'),
           os.path.join(CSRC, 'oa_pjoin_core.cuh'),
           os.path.join(os.path.dirname(HERE), 'include', 'orbit_b200.h')]

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',   # B200 only, no fat binary
    '-lineinfo', '-O3', '-std=c++17',
    '-fmad=false',            # numpy never contracts a*b+c; neither do we
    '-Xcompiler', '-fPIC', '-Xcompiler', '-O2',
]


def _nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found; cannot build liborbit_b200.so')
    return exe


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link the shared library."""
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
        if verbose:
            cmd.insert(1, '-Xptxas')
            cmd.insert(2, '-v')
            print(' '.join(cmd))
        subprocess.run(cmd, check=True)
        objs.append(obj)
    cmd = [nvcc, '-shared', '-o', LIB] + objs + \
        ['-gencode', 'arch=compute_100a,code=sm_100a']
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
