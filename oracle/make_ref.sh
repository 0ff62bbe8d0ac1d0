#!/bin/bash
# Installs the reference's OWN toy implementation as oracle/_ref/ (test infrastructure, git-ignored, travels to the GPU
# box with the gpurun snapshot): benchmark/integrators/numba_integrators.py is copied VERBATIM from the read-only
# reference checkout -- it imports only numpy and numba, both present in this image and on the GPU box.  The rest of the
# reference's toy path (benchmark/testsystems/low_dimensional_systems.py) imports simtk / openmmtools at module level
# and cannot be loaded; oracle/ref_toy.py restates its 25-line protocol loop around the copied closures.
# Nothing of the reference enters the repository history: oracle/_ref/ is listed in .gitignore.
#   usage: oracle/make_ref.sh [reference checkout, default /root/reference]
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
REF=${1:-/root/reference}
SRC=$REF/benchmark/integrators/numba_integrators.py
if [ ! -f "$SRC" ]; then echo "make_ref.sh: $SRC not found (no reference checkout here): keeping oracle/_ref as it is"; exit 0; fi
mkdir -p "$HERE/_ref"
cp "$SRC" "$HERE/_ref/numba_integrators.py"
sha256sum "$HERE/_ref/numba_integrators.py" | cut -d' ' -f1 > "$HERE/_ref/numba_integrators.sha256"
echo "oracle/_ref/numba_integrators.py <- $SRC ($(cat "$HERE/_ref/numba_integrators.sha256" | cut -c1-12))"
