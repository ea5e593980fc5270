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
           'oa_pjoin.cu', 'oa_pj2.cu', 'oa_regions.cu']
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


def build(force=False, verbose=False, extra_flags=(), out=None):
    """Compile every CUDA source for sm_100a and link the shared library.

    ``extra_flags`` / ``out``: tuning builds (e.g. ``-DOA_PJOIN_THREADS=256``)
    written next to the default library and selected with ``OA_LIB_PATH``."""
    if out is None and not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, 'build' if out is None else
                          'build_' + os.path.basename(out).replace('.so', ''))
    os.makedirs(objdir, exist_ok=True)
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + [
            '-c', os.path.join(CSRC, src), '-o', obj]
        if verbose:
            cmd.insert(1, '-Xptxas')
            cmd.insert(2, '-v')
            print(' '.join(cmd))
        subprocess.run(cmd, check=True)
        objs.append(obj)
    target = LIB if out is None else out
    cmd = [nvcc, '-shared', '-o', target] + objs + \
        ['-gencode', 'arch=compute_100a,code=sm_100a']
    subprocess.run(cmd, check=True)
    return target


if __name__ == '__main__':
    # python -m nbody_orbit_analysis_b200._build [--force] [-v]
    #        [--out variants/liborbit_b200_small.so -DOA_PJOIN_THREADS=256 ...]
    argv = sys.argv[1:]
    out_path = None
    if '--out' in argv:
        out_path = os.path.abspath(argv[argv.index('--out') + 1])
        os.makedirs(os.path.dirname(out_path), exist_ok=True)
    print(build(force='--force' in argv, verbose='-v' in argv,
                extra_flags=[f for f in argv if f.startswith('-D')],
                out=out_path))
