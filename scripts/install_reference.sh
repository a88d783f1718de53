#!/bin/bash
# Reference arm / drop-in tests: make the UNMODIFIED reference importable on the GPU box (which has no /root/reference).
# The reference has no setup.py / pyproject.toml (README installs a conda env), so `pip install --target baseline/_ref
# /root/reference` cannot work; what an install would do -- place the package tree -- is done by copying it into
# baseline/_ref/, which is git-ignored (never committed) but NOT gpurun-ignored, so it travels with the snapshot.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC=${1:-/root/reference}
[ -d "$SRC/footprints" ] || { echo "no reference checkout at $SRC"; exit 0; }
mkdir -p "$ROOT/baseline/_ref"
if python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$ROOT/baseline/_ref" "$SRC" >/dev/null 2>&1; then
  echo "pip install ok"
else
  rm -rf "$ROOT/baseline/_ref/footprints" "$ROOT/baseline/_ref/test_data"
  cp -r "$SRC/footprints" "$ROOT/baseline/_ref/footprints"
  cp -r "$SRC/test_data" "$ROOT/baseline/_ref/test_data"
  cp "$SRC/paths.yaml" "$ROOT/baseline/_ref/paths.yaml"
  echo "reference has no build metadata: package tree placed under baseline/_ref (git-ignored)"
fi
