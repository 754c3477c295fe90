"""Recipe that vendors the UNMODIFIED Python reference of the hot path into oracle/_ref/ (git-ignored, NOT gpurun-ignored:
it travels to the GPU box like a built .so) so that bench.py's reference arm and cpu_baseline time the REAL reference
(`kind: "reference"`) on the GPU box's host cores, where /root/reference does not exist.

The reference is pure Python (no build step): the "build" is a verbatim copy of the files the path imports -
src/nr/network/*.py, src/nr/utils/field_utils.py, src/nr/configs/nrvgn_sdf.yaml and src/gd/{__init__,networks}.py -
keeping the tree layout, so tests/golden/ref_harness.py works on it with GRASPNERF_REFERENCE=oracle/_ref.  Nothing under
oracle/_ref is committed; nothing in the product path (graspnerf_b200/) reads it.

usage: python oracle/make_ref.py [reference_root]      (default /root/reference; no-op when it is absent)
"""
import glob
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')
FILES = ['src/nr/network/*.py', 'src/nr/utils/field_utils.py', 'src/nr/configs/nrvgn_sdf.yaml',
         'src/gd/__init__.py', 'src/gd/networks.py']


def make_ref(ref_root='/root/reference', verbose=False):
    if not os.path.isdir(os.path.join(ref_root, 'src', 'nr', 'network')):
        return None
    n = 0
    for pat in FILES:
        for src in glob.glob(os.path.join(ref_root, pat)):
            rel = os.path.relpath(src, ref_root)
            dst = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            if not os.path.exists(dst) or open(src, 'rb').read() != open(dst, 'rb').read():
                shutil.copyfile(src, dst)
            n += 1
    if verbose:
        print(f'oracle/_ref: {n} reference files from {ref_root}', file=sys.stderr)
    return DST


def ref_available():
    return os.path.isdir(os.path.join(DST, 'src', 'nr', 'network'))


if __name__ == '__main__':
    make_ref(sys.argv[1] if len(sys.argv) > 1 else '/root/reference', verbose=True)
