#!/bin/bash
# Install the UNMODIFIED reference (mjhoptics/ray-optics, /root/reference) into baseline/_ref
# for bench.py's reference arm.  Build container only (needs /root/reference); the result is
# git-ignored and travels to the GPU box with the gpurun snapshot.
#   --no-deps: opticalglass / anytree / json_tricks / pyside6 ... are not in the offline
#   wheelhouse; baseline/reference_arm.py says how the hot path runs without them.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
[ -d /root/reference ] || { echo "no /root/reference here: nothing to install"; exit 0; }
rm -rf /tmp/refcopy && cp -r /root/reference /tmp/refcopy      # the build writes into the source tree
rm -rf "$HERE/_ref"
python -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$HERE/_ref" /tmp/refcopy
# package data that setuptools drops without its scm file finder (read at import time by
# rayoptics/util/colour_system.py and friends)
(cd /root/reference/src && find rayoptics -type f \( -name '*.txt' -o -name '*.csv' \) ! -path '*/tests/*' \
    -exec cp --parents {} "$HERE/_ref/" \;)
rm -rf /tmp/refcopy
echo "installed: $(ls "$HERE/_ref" | tr '\n' ' ')"
