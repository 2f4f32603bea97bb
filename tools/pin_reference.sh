#!/bin/sh
# One command to pin the oracle to the real reference the day SeqAn2 is available:
#
#   SEQAN_INCLUDE=/path/to/seqan/include tools/pin_reference.sh [configs]      (default configs: 1,3small,5)
#
# Builds oracle/_ref/talc from the sources where they lie under /root/reference/src (recipe: oracle/Makefile `ref`,
# the flags of the reference's Makefile:3; nothing is copied into this repository), or uses baseline/_ref/talc if a
# driver-installed binary exists, then diffs .fa / sorted .log / .config.txt of reference vs oracle on the synthetic
# configurations and against tests/golden/oracle_sha256.json (tools/pin_reference.py).
set -e
cd "$(dirname "$0")/.."
REF=""
if [ -x baseline/_ref/talc ]; then REF=baseline/_ref/talc; fi
if [ -z "$REF" ]; then
  if [ -z "$SEQAN_INCLUDE" ] || [ ! -d "$SEQAN_INCLUDE/seqan" ]; then
    echo "pin_reference: no baseline/_ref/talc and SEQAN_INCLUDE does not point at a SeqAn2 include tree -> parity stays UNPINNED" >&2
    exit 2
  fi
  make -C oracle ref SEQAN_INCLUDE="$SEQAN_INCLUDE"
  REF=oracle/_ref/talc
fi
exec python tools/pin_reference.py --ref "$REF" --configs "${1:-1,3small,5}"
