#!/bin/bash
# Install the UNMODIFIED reference into baseline/_ref (git-ignored; it travels to the GPU box with gpurun):
#   * `pip install --no-index --no-build-isolation --no-deps --target baseline/_ref` of a scratch copy of /root/reference
#     (setup.py = find_packages(): that installs `lit_gpt`; --no-deps because `lightning @ git+...` cannot be resolved offline);
#   * the reference's `generate/` and `quantize/` are script directories without __init__.py, so pip does not package them: they are
#     copied next to the package (namespace packages, exactly how the reference imports them from its repo root);
#   * oracle/_shims (import stubs for the absent lightning / lightning_utilities / nltk) go to baseline/_ref/_shims.
# Only bench.py's reference arms import from here.  No file of the reference is modified.
set -e
REPO="$(cd "$(dirname "$0")/.." && pwd)"
REF=${1:-/root/reference}
[ -d "$REF" ] || { echo "no reference tree at $REF: keeping whatever baseline/_ref holds"; exit 0; }
rm -rf /tmp/_lp_refcopy "$REPO/baseline/_ref"
cp -r "$REF" /tmp/_lp_refcopy
(cd /tmp/_lp_refcopy && python -m pip install -q --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps \
    --target "$REPO/baseline/_ref" . ) || { echo "pip install failed: copying the package instead"; mkdir -p "$REPO/baseline/_ref"; cp -r "$REF/lit_gpt" "$REPO/baseline/_ref/"; }
rm -rf "$REPO/baseline/_ref/s2l"
cp -r "$REF/generate" "$REF/quantize" "$REPO/baseline/_ref/"
cp -r "$REPO/oracle/_shims" "$REPO/baseline/_ref/_shims"
rm -rf /tmp/_lp_refcopy
ls "$REPO/baseline/_ref"
