#!/bin/bash
# The one offline install of the task: the UNMODIFIED reference into baseline/_ref (git-ignored; it travels to the GPU box with
# the gpurun snapshot), so that bench.py can time the real Python engine on the GPU box's host cores in the same run
# (cpu_baseline.python_reference_same_box).  The reference ships neither setup.py nor pyproject.toml, so
#   python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference
# fails ("neither 'setup.py' nor 'pyproject.toml' found"); as the task allows, the install is made from a copy under /tmp to
# which ONLY a setup.py naming the reference's top-level modules is added -- the .py sources are installed byte for byte
# (checked below with cmp).  Runs only where /root/reference exists (the build container).
set -euo pipefail
REF=${1:-/root/reference}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
[ -d "$REF" ] || { echo "no reference tree at $REF"; exit 0; }
TMP=$(mktemp -d /tmp/tarok_ref_XXXX)
cp "$REF"/*.py "$TMP"/
MODS=$(cd "$REF" && ls *.py | sed 's/\.py$//' | tr '\n' ' ')
python - "$TMP" $MODS <<'PY'
import sys
tmp, mods = sys.argv[1], sys.argv[2:]
open(tmp + "/setup.py", "w").write(
    "from setuptools import setup\nsetup(name='anzeA-Tarok-reference', version='0', py_modules=%r)\n" % mods)
PY
rm -rf "$ROOT/baseline/_ref"
mkdir -p "$ROOT/baseline"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$ROOT/baseline/_ref" "$TMP" > "$TMP/pip.log" 2>&1 \
  || { tail -5 "$TMP/pip.log"; exit 1; }
for f in "$REF"/*.py; do cmp -s "$f" "$ROOT/baseline/_ref/$(basename "$f")" || { echo "DIFFERS: $f"; exit 1; }; done
echo "installed $(ls "$ROOT/baseline/_ref"/*.py | wc -l) unmodified modules into baseline/_ref"
rm -rf "$TMP"
