#!/bin/sh
# Copies the reference checkout into baseline/_ref/reference (git-ignored, NOT gpurun-ignored) so that
# tests/test_dropin_trainer.py can run the UNMODIFIED reference trainer on the GPU box, where /root/reference does
# not exist.  The copy never enters the repo's history.
set -e
here=$(cd "$(dirname "$0")/.." && pwd)
src=${1:-/root/reference}
mkdir -p "$here/baseline/_ref"
rm -rf "$here/baseline/_ref/reference"
cp -r "$src" "$here/baseline/_ref/reference"
rm -rf "$here/baseline/_ref/reference/.git" "$here/baseline/_ref/reference/figs"
echo "staged $src -> $here/baseline/_ref/reference"
