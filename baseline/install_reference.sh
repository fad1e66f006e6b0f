#!/bin/bash
# Install the UNMODIFIED reference package into baseline/_ref (git-ignored; it travels to the GPU box with the snapshot).
# The reference ships no setup.py / pyproject, so the install runs from a copy under /tmp that adds a three-line setup.py
# next to the package (the package files themselves are byte-identical: `diff -r /root/reference/movedepth
# baseline/_ref/movedepth` shows only __pycache__ and the data files left out below).  Only the split lists the Trainer
# constructor reads (splits/eigen_zhou/*.txt, trainer.py:159-161) are packaged; the 57 MB of KITTI ground-truth depths
# and the ImageNet checkpoints are not needed (`--weights_init scratch`, synthetic inputs).
set -e
REPO="$(cd "$(dirname "$0")/.." && pwd)"
REF="${MOVEDEPTH_REFERENCE:-/root/reference}"
TMP="$(mktemp -d /tmp/movedepth_ref.XXXXXX)"
cp -r "$REF/movedepth" "$TMP/"
cat > "$TMP/setup.py" <<'PY'
from setuptools import setup
setup(name="movedepth", version="0.0.0", packages=["movedepth", "movedepth.networks", "movedepth.datasets"],
      package_data={"movedepth": ["splits/eigen_zhou/*.txt"]})
PY
rm -rf "$REPO/baseline/_ref"
cd "$TMP" && python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps --target "$REPO/baseline/_ref" "$TMP"
rm -rf "$TMP"
