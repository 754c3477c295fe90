"""Builds graspnerf_b200/lib/libgraspnerf_b200.so with nvcc for sm_100a (in-tree, so the .so travels to the GPU box).

The library has no torch dependency: plain CUDA runtime + extern "C" entry points (include/graspnerf_b200.h).
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
LIBDIR = os.path.join(PKG, 'lib')
LIBPATH = os.path.join(LIBDIR, 'libgraspnerf_b200.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest():
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    files.append(os.path.join(os.path.dirname(PKG), 'include', 'graspnerf_b200.h'))
    for f in files:
        if os.path.isfile(f):
            h.update(f.encode())
            with open(f, 'rb') as fh:
                h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return 'nvcc'


def build_library(force=False, verbose=True, tag='', defs=()):
    """tag / defs: development builds of kernel variants (tools/ab_build.py): lib/libgraspnerf_b200<tag>.so compiled with
    extra -D flags, loaded instead of the product library when GN_LIB_TAG=<tag> is set (A/B timing in one GPU call)."""
    os.makedirs(LIBDIR, exist_ok=True)
    libpath = LIBPATH[:-3] + tag + '.so'
    stamp = os.path.join(LIBDIR, 'build%s.sha256' % tag)
    dig = _digest() + ' ' + ' '.join(defs)
    if not force and os.path.exists(libpath) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return libpath
    nvcc = nvcc_path()
    objs = []

    def compile_one(src):
        obj = os.path.join(LIBDIR, src[:-3] + tag + '.o')
        cmd = [nvcc] + NVCC_FLAGS + list(defs) + ['-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
        return obj
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [nvcc, '-shared', '--cudart', 'shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', libpath] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    with open(stamp, 'w') as f:
        f.write(dig)
    if verbose:
        print('built', libpath, file=sys.stderr)
    return libpath


if __name__ == '__main__':
    build_library(force='--force' in sys.argv)
