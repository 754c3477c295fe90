"""The reference import harness lives in oracle/ref_harness.py (shared with bench.py's CPU legs); re-exported here for
tests/golden/make_golden.py and the live-reference tests."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.ref_harness import *          # noqa: F401,F403,E402
from oracle.ref_harness import REF_ROOT, reference_available, install_shims, load_reference, build_reference_net, shims  # noqa: F401,E402
