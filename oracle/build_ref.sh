#!/bin/sh
# TEST / BENCH INFRASTRUCTURE ONLY.
# Ships the UNMODIFIED reference to the GPU box: copies /root/reference (src/, configs/, data/, licences; ~1 MB of Python plus the
# two StyleGAN2 CUDA ops) into the git-ignored directory oracle/_ref/reference/.  oracle/_ref/ is NOT gpurun-ignored, so the copy
# travels with the snapshot exactly like a built .so; it never enters the git history.  Nothing is edited: oracle/ref_import.py makes
# the tree importable through in-memory shims and `bench.py --impl reference` / the `incumbent_gpu` block
# run the reference's own modules (CPU: kind "reference"; cuda:0: the incumbent-GPU bar).
# Run in the build container (where /root/reference is mounted):  sh oracle/build_ref.sh
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${GA_REFERENCE_SRC:-/root/reference}"
DST="$HERE/_ref/reference"
if [ ! -d "$SRC/src/defenses/ours" ]; then
  echo "build_ref.sh: no reference tree at $SRC (nothing to do)"; exit 0
fi
rm -rf "$DST"
mkdir -p "$DST"
cp -r "$SRC/src" "$SRC/configs" "$SRC/data" "$DST/"
for f in "$SRC"/LICENSE* "$SRC"/NVAE_LICENSE "$SRC"/README.md; do [ -f "$f" ] && cp "$f" "$DST/"; done
find "$DST" -name __pycache__ -type d -prune -exec rm -rf {} + 2>/dev/null || true
( cd "$SRC" && find src configs -type f | LC_ALL=C sort | xargs sha256sum ) > "$HERE/_ref/reference.sha256"
( cd "$DST" && find src configs -type f | LC_ALL=C sort | xargs sha256sum ) | cmp -s - "$HERE/_ref/reference.sha256" \
  || { echo "build_ref.sh: copy differs from the source tree"; exit 1; }
echo "build_ref.sh: $(find "$DST" -type f | wc -l) files of the unmodified reference -> $DST"
