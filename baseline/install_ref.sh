#!/bin/bash
# Installs the UNMODIFIED reference into baseline/_ref (git-ignored, travels to the GPU box with gpurun).
#   * `pip install --target` of /root/reference (its setup.py packages trajnetbaselines + trajnetbaselines.lstm);
#     built from a copy under /tmp because /root/reference is read-only; --no-deps: pykalman, pysparkling,
#     trajnetplusplustools, python-json-logger and torch==1.10.0 are not in the offline wheelhouse.
#   * the sub-packages sgan / classical / vae that trajnetbaselines/__init__.py imports but setup.py does not
#     list, and the reference's top-level `evaluator/` package, which trajnetbaselines/lstm/trajnet_evaluator.py imports
#     but setup.py does not list, is copied beside it (same unmodified files).
# Nothing under baseline/_ref is product source; bench.py --impl reference and tests/test_dropin.py load it through
# oracle/ref_shim.py (stub modules for the absent third-party imports).
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${TRAJNET_REFERENCE_SRC:-/root/reference}"
[ -d "$SRC/trajnetbaselines" ] || { echo "reference sources not found at $SRC"; exit 1; }
rm -rf /tmp/_trajnet_ref_src "$ROOT/baseline/_ref"
cp -r "$SRC" /tmp/_trajnet_ref_src
( cd /tmp/_trajnet_ref_src && python -m pip install --no-index --no-build-isolation --no-deps \
    --find-links /opt/wheelhouse --target "$ROOT/baseline/_ref" /tmp/_trajnet_ref_src )
cp -r "$SRC/evaluator" "$ROOT/baseline/_ref/evaluator"
# setup.py lists only trajnetbaselines and trajnetbaselines.lstm, but trajnetbaselines/__init__.py imports
# .sgan, .classical and .vae: the wheel alone is not importable.  The same unmodified sub-packages go beside it.
for sub in sgan classical vae; do
  [ -d "$ROOT/baseline/_ref/trajnetbaselines/$sub" ] || cp -r "$SRC/trajnetbaselines/$sub" "$ROOT/baseline/_ref/trajnetbaselines/$sub"
done
rm -rf /tmp/_trajnet_ref_src
find "$ROOT/baseline/_ref" -name __pycache__ -prune -exec rm -rf {} +
echo "installed: $(ls "$ROOT/baseline/_ref")"
