#!/bin/bash
# Installs the UNMODIFIED reference (GoekeLab/m6anet, /root/reference) into oracle/_ref/ with pip.
# TEST / BENCH INFRASTRUCTURE (see oracle/__init__.py): oracle/_ref is git-ignored -- no reference source enters the
# history -- but it is not gpurun-ignored, so it travels to the GPU box like the built .so files, where
# `bench.py --impl reference` times the reference's own CPU code (kind "reference") and tests/test_ref_arm.py
# checks the oracle against it.  Needs /root/reference (build container only); a no-op success when it is absent
# and oracle/_ref already exists.
#   --ignore-requires-python : the reference's setup.py pins python <3.9; its code runs unchanged on 3.12
#   --no-deps                : its pins (torch==1.6.0, ...) are not installable; the image's torch/numpy/pandas are used
#   ujson shim               : the only missing import (m6anet/__init__.py -> scripts/dataprep.py -> ujson), json-compatible
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${M6A_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/m6anet" ]; then
  if [ -d "$OUT/m6anet" ]; then echo "make_ref: $REF absent, keeping existing $OUT"; exit 0; fi
  echo "make_ref: $REF absent and $OUT missing" >&2; exit 1
fi
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
cp -r "$REF" "$TMP/src"            # the reference tree is read-only; setup.py writes build/ and egg-info next to itself
rm -rf "$OUT"
python -m pip install --quiet --no-index --no-build-isolation --no-deps --ignore-requires-python \
    --find-links /opt/wheelhouse --target "$OUT" "$TMP/src"
printf '"""json-compatible stand-in for the one dependency of the reference that this image lacks."""\nfrom json import *  # noqa: F401,F403\n' > "$OUT/ujson.py"
( cd "$REF" && find m6anet -name '*.py' -not -path 'm6anet/tests/*' -print0 | sort -z | xargs -0 sha256sum ) > "$OUT/SOURCES.sha256"
( cd "$OUT" && sha256sum -c --quiet SOURCES.sha256 )     # installed == reference, file by file
echo "make_ref: installed $(python - <<PY
import sys; sys.path.insert(0, "$OUT")
import m6anet; print("m6anet", m6anet.__version__)
PY
) into $OUT"
