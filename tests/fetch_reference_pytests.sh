#!/bin/bash
# Copy the reference's own Python tests for the hot path, unmodified, into a git-ignored scratch
# directory so that tests/test_gpu_reference_pytests.py can run them on the GPU box (which has no
# /root/reference).  Nothing here is committed.
set -e
SRC=${1:-/root/reference/lightmotif-py/lightmotif/tests}
DST=$(dirname "$0")/_ref_pytests
mkdir -p "$DST"
for f in test_pipeline.py test_scanner.py test_sequence.py; do cp "$SRC/$f" "$DST/$f"; done
ls -la "$DST"
