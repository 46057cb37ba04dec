#!/bin/bash
# The -DLJMD_DEBUG_CHECKS build (in-kernel bounds assertions on every staged window, list index and bulk
# copy of the cell path) under the whole GPU test-suite plus a 4M-particle run with rebuilds.  This is the
# stand-in for compute-sanitizer memcheck, which is closed on the measurement pool
# (profiles/r2/sanitizer_closed_message.log); races would show up in the bit-exact determinism tests.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd); cd "$ROOT"
scripts/build_variant.sh debug -DLJMD_DEBUG_CHECKS > /dev/null
export LJMD_LIB=$ROOT/variants/debug/libljmd.so
python -m pytest tests -m gpu -q 2>&1 | tail -3
python scripts/cells_one.py 4194304 60 0.5 200
