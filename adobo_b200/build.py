"""Builds adobo_b200/libadobo_b200.so (sm_100a only) with nvcc, in-tree.

    python -m adobo_b200.build [--force]

nvcc cross-compiles without a GPU.  The library links cudart statically and resolves NCCL with
dlopen at run time, so importing it needs nothing but the CUDA driver.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
# ADOBO_B200_LIB_SUFFIX: diagnostic variants (e.g. _tl with ADOBO_B200_CFLAGS=-DADOBO_B200_TIMELINE) beside the product
SUFFIX = os.environ.get('ADOBO_B200_LIB_SUFFIX', '')
OUT = os.path.join(HERE, 'libadobo_b200%s.so' % SUFFIX)
SOURCES = ['api.cu', 'normalize.cu', 'moments.cu', 'gather.cu', 'irlb.cu', 'svd_arrow.cu', 'knn.cu', 'knn_tc.cu', 'snn.cu', 'synth.cu']
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
FLAGS = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr',
         '-Xptxas', '-v']


def _nvcc():
    for cand in (shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found')


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, '..', 'include', 'adobo_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return OUT
    nvcc = _nvcc()
    objdir = os.path.join(HERE, 'build' + SUFFIX)
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace('.cu', '.o'))
        cmd = [nvcc] + ARCH + FLAGS + os.environ.get('ADOBO_B200_CFLAGS', '').split() + ['-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj + '.log', 'w') as f:
            f.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed on %s:\n%s' % (src, r.stderr[-4000:]))
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    # link beside the target and rename: the library on disk is always a complete file (a snapshot of the tree may be
    # taken for a GPU run at any moment)
    cmd = [nvcc] + ARCH + ['-shared', '-o', OUT + '.tmp'] + objs + ['-ldl']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stderr[-4000:])
    os.replace(OUT + '.tmp', OUT)
    if verbose:
        print('built', OUT)
    return OUT


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose=True)
