"""Recipe for oracle/_ref — the UNMODIFIED reference package, installed next to the oracle (TEST INFRASTRUCTURE).

    python oracle/build_ref.py            # build container only: needs /root/reference

`pip install --no-index --no-build-isolation --no-deps --target oracle/_ref <copy of /root/reference>` (the source tree is
read-only, so the wheel is built from a copy under a temporary directory).  oracle/_ref/ is git-ignored (no reference
source enters the history) but NOT gpurun-ignored: it travels to the GPU box like the built .so, where
`bench.py --impl reference` and the `cpu_baseline` leg drive the reference's own `DeepFM` class through
`oracle.ref_loader` (cpu_baseline.kind = "reference").  When oracle/_ref is absent those legs fall back to the oracle port
(oracle/restatement.py, kind = "port").  Nothing under rec_pangu_b200/ may import this.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = '/root/reference'
TARGET = os.path.join(HERE, '_ref')


def build(verbose=True):
    """Returns the install directory, or None when the reference source tree is not present (GPU box)."""
    if not os.path.isdir(REF_SRC):
        return TARGET if os.path.isdir(os.path.join(TARGET, 'rec_pangu')) else None
    stamp = os.path.join(TARGET, '.installed_from')
    if os.path.isdir(os.path.join(TARGET, 'rec_pangu')) and os.path.exists(stamp):
        return TARGET
    tmp = tempfile.mkdtemp(prefix='rpb_ref_')
    try:
        src = os.path.join(tmp, 'reference')
        shutil.copytree(REF_SRC, src, ignore=shutil.ignore_patterns('.git', '*.pth', '*.csv', '__pycache__'))
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        cmd = [sys.executable, '-m', 'pip', 'install', '--no-index', '--no-build-isolation', '--no-deps', '--quiet',
               '--find-links', '/opt/wheelhouse', '--target', TARGET, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('pip install of the reference failed:\n' + r.stdout + r.stderr)
        with open(stamp, 'w') as f:
            f.write(REF_SRC + '\n')
        if verbose:
            print('[oracle.build_ref] installed the reference package into', TARGET)
        return TARGET
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == '__main__':
    print(build())
