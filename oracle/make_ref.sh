#!/usr/bin/env bash
# oracle/_ref: the UNMODIFIED Python reference (tachyon-beep/townlet), where it can travel to the GPU box.
#
# The reference is pure Python, so there is nothing to compile: this recipe copies the three trees the hot path
# needs (src/, configs/, scripts/, tests/data/observations/) from the read-only checkout into the git-ignored oracle/_ref/ (listed in
# .gitignore, NOT in .gpurunignore, so it ships with the gpurun snapshot like the built .so files).  It is test
# infrastructure: bench.py's cpu_baseline leg times it on the box's host cores and the composed drop-in test runs
# the drop-in classes beside it; the product package never imports it.  No reference source enters the history.
#
#   oracle/make_ref.sh [/path/to/reference]      (default /root/reference)
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
ref="${1:-/root/reference}"
if [ ! -d "$ref/src/townlet" ]; then
    echo "make_ref: no reference checkout at $ref (nothing to do)" >&2
    exit 0
fi
rm -rf "$here/_ref"
mkdir -p "$here/_ref"
for tree in src configs scripts; do
    cp -r "$ref/$tree" "$here/_ref/$tree"
done
cp "$ref/pyproject.toml" "$here/_ref/pyproject.toml"
# the reference's own stored observation goldens (four small .npz) for the composed drop-in test
mkdir -p "$here/_ref/tests/data"
cp -r "$ref/tests/data/observations" "$here/_ref/tests/data/observations"
chmod -R u+w "$here/_ref"
find "$here/_ref" -name __pycache__ -type d -prune -exec rm -rf {} +
git -C "$ref" rev-parse HEAD > "$here/_ref/REVISION" 2>/dev/null || echo "unknown" > "$here/_ref/REVISION"
echo "make_ref: copied $(find "$here/_ref" -type f | wc -l) files from $ref"
