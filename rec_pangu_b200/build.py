"""Build recipe for librec_pangu_b200.so (sm_100a only, in-tree so the .so travels with the repo snapshot).

    python -m rec_pangu_b200.build            # incremental
    python -m rec_pangu_b200.build --force
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'build')
LIB = os.path.join(HERE, 'librec_pangu_b200.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC,-fvisibility=hidden', '--threads', '2']


def _nvcc():
    for c in (shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if c and os.path.exists(c):
            return c
    raise RuntimeError('nvcc not found: librec_pangu_b200.so cannot be built')


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    hdrs.append(os.path.join(os.path.dirname(HERE), 'include', 'rec_pangu_b200.h'))
    return max(os.path.getmtime(h) for h in hdrs)


def build(force=False, verbose=True):
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    hdr_m = _deps_mtime()
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_m):
            jobs.append([nvcc] + NVCC_FLAGS + ['-c', src, '-o', obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed: ' + ' '.join(cmd) + '\n' + r.stdout + r.stderr)
        return cmd[-3]

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for s in ex.map(run, jobs):
                if verbose:
                    print('[rec_pangu_b200.build] compiled', os.path.basename(s))
    if jobs or not os.path.exists(LIB):
        run([nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB] + objs)
        if verbose:
            print('[rec_pangu_b200.build] linked', LIB)
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv)
