"""python -m graspnerf_b200 <reference script> [args...]: runs an UNMODIFIED reference entry script (src/nr/run_training.py,
scripts/sim_grasp.py) with the CUDA hot path installed in the reference's model registry (graspnerf_b200.install).
Run it from the reference checkout's root, exactly where train.sh / run_simgrasp.sh run their python commands."""
import os
import runpy
import sys


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    script = sys.argv[1]
    root = os.getcwd()
    for p in (os.path.join(root, 'src', 'nr'), os.path.join(root, 'src'), os.path.dirname(os.path.abspath(script))):
        if os.path.isdir(p) and p not in sys.path:
            sys.path.insert(0, p)
    from . import install
    install(verbose=True)
    sys.argv = sys.argv[1:]
    runpy.run_path(script, run_name='__main__')


if __name__ == '__main__':
    main()
