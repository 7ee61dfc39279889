#!/bin/bash
# Installs the UNMODIFIED reference (inverted-ai/torchdrivesim 0.2.3) into baseline/_ref/ (git-ignored, but it
# travels to the GPU box with the repository snapshot).  The reference's source tree is read-only, so pip builds
# from a copy under /tmp; its dependencies pandas / omegaconf / shapely / lanelet2 / ... are not in the offline
# wheelhouse, hence --no-deps (baseline/ref_import.py stubs the three that are imported at module level).  The
# wheel leaves out the bundled maps (no package_data in the reference's setup), so the data files of
# torchdrivesim/resources are copied next to the installed package: find_map_config() looks there.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${1:-/root/reference}"
[ -d "$SRC/torchdrivesim" ] || { echo "no reference under $SRC"; exit 1; }
rm -rf /tmp/tds_refcopy "$HERE/_ref"
cp -r "$SRC" /tmp/tds_refcopy
python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target "$HERE/_ref" /tmp/tds_refcopy \
  || python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$HERE/_ref" /tmp/tds_refcopy
mkdir -p "$HERE/_ref/torchdrivesim/resources"
cp -r "$SRC/torchdrivesim/resources/maps" "$HERE/_ref/torchdrivesim/resources/"
chmod -R u+w "$HERE/_ref"
rm -rf /tmp/tds_refcopy
echo "installed: $(ls "$HERE/_ref")"
