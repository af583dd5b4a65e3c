#!/bin/sh
# Builds libqgt_b200.so (+ the C compatibility library and the oracle) in-tree for sm_100a: the same steps
# __graft_entry__.build() runs.  Extra arguments: --force rebuilds everything.
set -e
cd "$(dirname "$0")"
python -c "import sys, __graft_entry__ as g; g.build(force='--force' in sys.argv)" "$@"
