#!/bin/bash
# Stages the UNMODIFIED reference checkout under baseline/_ref/reference (git-ignored, NOT gpurun-ignored) so that the
# "scripts run unchanged" GPU tests (tests/test_gpu_dropin_scripts.py) can execute the reference's own script files on
# the GPU box, where /root/reference does not exist.  Nothing under baseline/_ref/ is ever committed.
set -e
SRC=${1:-/root/reference}
DST="$(dirname "$0")/../baseline/_ref/reference"
rm -rf "$DST"; mkdir -p "$DST"
cp -r "$SRC"/. "$DST"/
rm -rf "$DST/.git"
echo "staged $(find "$DST" -name '*.py' | wc -l) python files under $DST"
