"""Development aid: builds tagged variants of the library with extra -D flags for A/B timing in one GPU call.
usage: python tools/ab_build.py _tag=-DFOO,-DBAR=1 [...]   ->  graspnerf_b200/lib/libgraspnerf_b200_tag.so   (GN_LIB_TAG=_tag)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graspnerf_b200.build import build_library

for a in sys.argv[1:]:
    tag, defs = a.split('=', 1)
    print(build_library(tag=tag, defs=tuple(d for d in defs.split(',') if d)))
