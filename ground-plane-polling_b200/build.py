"""Build recipe of libgpp.so (hand-written CUDA for sm_100a, C ABI in include/gpp.h).

    python ground-plane-polling_b200/build.py [--force]

nvcc cross-compiles without a GPU.  The .so is written next to this file (git-ignored; it travels to the
GPU box with the repo snapshot) so that the process that loads it shows an in-tree native library.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'build')
LIB = os.path.join(HERE, 'libgpp.so')
SOURCES = ['gpp_api.cu', 'gpp_launch.cu', 'gpp_pose.cu', 'gpp_detect.cu', 'gpp_microbench.cu', 'gpp_order.cu']
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
NVCC_FLAGS = ARCH + ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return 'nvcc'


def _deps():
    out = [os.path.join(ROOT, 'include', 'gpp.h')]
    for f in os.listdir(CSRC):
        out.append(os.path.join(CSRC, f))
    return out


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def _compile(src):
    obj = os.path.join(OBJ, src.replace('.cu', '.o'))
    extra = os.environ.get('GPP_NVCC_EXTRA', '').split()          # experiments only, e.g. -DGPP_M6_UNROLL=2
    cmd = [_nvcc()] + NVCC_FLAGS + extra + ['-c', os.path.join(CSRC, src), '-o', obj]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
    return src, obj, p.returncode, p.stdout


def build_libgpp(force=False, verbose=False, out=None):
    """Compile every CUDA source for sm_100a and link libgpp.so.  Returns the library path.
    ``out`` (experiments only) links to another file name, e.g. a variant built with GPP_NVCC_EXTRA."""
    global LIB, OBJ
    if out:
        LIB = os.path.join(HERE, out)
        OBJ = os.path.join(HERE, 'build', out.replace('.so', ''))
        force = True
    if not force and not is_stale():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    objs = []
    log = []
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        for src, obj, rc, out in ex.map(_compile, SOURCES):
            log.append('== %s ==\n%s' % (src, out))
            if rc != 0:
                raise RuntimeError('nvcc failed on %s:\n%s' % (src, out))
            objs.append(obj)
    with open(os.path.join(OBJ, 'ptxas.log'), 'w') as f:
        f.write('\n'.join(log))
    if verbose:
        sys.stdout.write('\n'.join(log) + '\n')
    cmd = [_nvcc()] + ARCH + ['-shared', '-o', LIB] + objs
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
    if p.returncode != 0:
        raise RuntimeError('link failed:\n%s' % p.stdout)
    return LIB


if __name__ == '__main__':
    _out = sys.argv[sys.argv.index('--out') + 1] if '--out' in sys.argv else None
    print(build_libgpp(force='--force' in sys.argv, verbose='-v' in sys.argv, out=_out))
