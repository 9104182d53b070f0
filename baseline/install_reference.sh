#!/bin/bash
# Installs the UNMODIFIED reference (ada-shen/Interpret_quality) into baseline/_ref/ for `bench.py --impl reference`.
# The reference ships no packaging metadata (no setup.py / pyproject.toml), so the stock command
#   python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference
# stops with "neither 'setup.py' nor 'pyproject.toml' found".  This script installs from a copy under /tmp that adds
# ONLY a setup.py naming the reference's modules (models/ and tools/ have no __init__.py: they are namespace packages
# and are listed as such); no reference source file is edited.  baseline/_ref/ is git-ignored (reference sources
# never enter the history) but travels to the GPU box with gpurun.
set -eu
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${INTERPRET_QUALITY_REF_SRC:-/root/reference}"
[ -d "$REF" ] || { echo "reference not found at $REF" >&2; exit 1; }
python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target "$HERE/_ref" "$REF" \
    > "$HERE/install_stock.log" 2>&1 && { echo "stock install worked"; exit 0; } || tail -2 "$HERE/install_stock.log"
TMP="$(mktemp -d)"
cp -r "$REF"/. "$TMP/"
rm -rf "$TMP/.git"
cat > "$TMP/setup.py" <<'PY'
import glob, os
from setuptools import setup
mods = [os.path.splitext(f)[0] for f in glob.glob("*.py") if f != "setup.py"]
setup(name="interpret_quality_reference", version="0", py_modules=mods, packages=["models", "tools"],
      package_data={}, zip_safe=False)
PY
rm -rf "$HERE/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$HERE/_ref" "$TMP" \
    > "$HERE/install_tmp.log" 2>&1 || { tail -20 "$HERE/install_tmp.log"; exit 1; }
rm -rf "$TMP"
ls "$HERE/_ref" | head -40
