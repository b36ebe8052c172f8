#!/bin/sh
# Installs the UNMODIFIED reference (mggg/GerryChain, pure Python) into oracle/_ref/ so that
# bench.py's CPU arm can time the reference's own ReCom on the GPU box's host cores.
#
#   sh oracle/make_ref.sh          # build container only: needs /root/reference
#
# oracle/_ref/ is git-ignored (no reference source enters the history) but NOT
# gpurun-ignored: like the built .so files it travels with the working tree.  The install
# is the reference's own setup.py through pip (offline, no dependency resolution: networkx,
# pandas and scipy are part of the image; geopandas is only a type hint there and is
# satisfied by the one-class stub oracle/refshim/geopandas, which bench.py puts on the path).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${GERRYCHAIN_REFERENCE:-/root/reference}"
[ -d "$REF/gerrychain" ] || { echo "make_ref.sh: $REF not found (GPU box: use the prebuilt oracle/_ref)"; exit 0; }
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
cp -r "$REF" "$TMP/src"          # /root/reference is read-only; setup.py writes egg-info
rm -rf "$HERE/_ref"
python -m pip install --quiet --no-index --no-build-isolation --no-deps \
    --find-links /opt/wheelhouse --target "$HERE/_ref" "$TMP/src"
python - <<PY
import sys
sys.path[:0] = ["$HERE/_ref", "$HERE/refshim"]
import gerrychain
print("oracle/_ref: gerrychain", gerrychain.__version__, "installed from $REF")
PY
