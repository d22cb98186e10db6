"""Build libcaspr_b200.so (hand-written sm_100a CUDA behind a C ABI) in-tree with nvcc.

    python -m caspr_b200.build [--force]

Each ``csrc/*.cu`` is compiled to an object with
``-gencode arch=compute_100a,code=sm_100a -lineinfo`` and linked into
``caspr_b200/lib/libcaspr_b200.so``.  nvcc cross-compiles without a GPU, so this also runs in
the development container.  The library is git-ignored but travels to the GPU box with the
repository snapshot.
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, 'lib')
OBJ_DIR = os.path.join(HERE, 'build')
LIB_PATH = os.path.join(LIB_DIR, 'libcaspr_b200.so')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')
EXPORTS = os.path.join(CSRC, 'exports.map')      # only the caspr_* C entry points leave the library

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-I', INCLUDE,
              '--expt-relaxed-constexpr']


def _nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found: libcaspr_b200.so cannot be built')
    return exe


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, INCLUDE):
        for f in sorted(os.listdir(root)):
            if f.endswith(('.cu', '.cuh', '.h', '.map')):
                with open(os.path.join(root, f), 'rb') as fh:
                    h.update(f.encode())
                    h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(src):
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + '.o')
    cmd = [_nvcc()] + NVCC_FLAGS + ['-c', src, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
    return obj


def build_library(force=False, verbose=False):
    """Compile (if sources changed) and return the path of libcaspr_b200.so."""
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp = os.path.join(LIB_DIR, 'libcaspr_b200.digest')
    digest = _digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == digest:
                return LIB_PATH
    srcs = _sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(_compile, srcs))
    cmd = [_nvcc(), '-shared', '-o', LIB_PATH] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a',
                                                         '-Xlinker', '--version-script=' + EXPORTS]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    with open(stamp, 'w') as f:
        f.write(digest)
    if verbose:
        print('built', LIB_PATH)
    return LIB_PATH


if __name__ == '__main__':
    print(build_library(force='--force' in sys.argv, verbose=True))
