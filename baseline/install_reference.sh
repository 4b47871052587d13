#!/bin/bash
# Installs the UNMODIFIED reference (albertogp71/qLDPCsim, read-only under /root/reference) into baseline/_ref with pip, from a
# copy under /tmp because the build writes into the source tree.  baseline/_ref is git-ignored (no reference source enters the
# history) but not gpurun-ignored, so it travels to the GPU box, where bench.py's reference arm and cpu_baseline leg time it.
#   --no-deps                  : stim is not installable here (no network); the decoders need NumPy only
#   --ignore-requires-python   : pyproject.toml asks for >= 3.13, the image has 3.12; nothing newer than 3.12 is used
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${1:-/root/reference}"
[ -f "$SRC/qLDPCsim/decoders.py" ] || { echo "reference tree not found under $SRC"; exit 1; }
TMP="$(mktemp -d)"
cp -r "$SRC" "$TMP/ref"
rm -rf "$HERE/_ref"
python -m pip install -q --no-index --no-build-isolation --no-deps --ignore-requires-python --find-links /opt/wheelhouse \
    --target "$HERE/_ref" "$TMP/ref"
rm -rf "$TMP"
cmp "$SRC/qLDPCsim/decoders.py" "$HERE/_ref/qLDPCsim/decoders.py" && cmp "$SRC/qLDPCsim/gf2math.py" "$HERE/_ref/qLDPCsim/gf2math.py"
echo "installed: $HERE/_ref/qLDPCsim"
