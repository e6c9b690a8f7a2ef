#!/usr/bin/env bash
# Ship the UNMODIFIED reference loss-head modules to the GPU box.
#
# /root/reference does not exist on the GPU box, so the three files the hot path needs
# (src/clip-event/model_clip.py, model_ot.py and the utils_image.py that model_clip imports) are
# copied byte for byte into oracle/_ref/ -- git-ignored (reference sources never enter the history)
# but not gpurun-ignored, so the copy travels with the snapshot.  bench.py's CPU legs
# (`--impl reference`, `cpu_baseline`) and tests import them from there through oracle/ref_loader.py.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="${1:-/root/reference/src/clip-event}"
DST="$HERE/_ref"
if [ ! -d "$SRC" ]; then
  echo "make_ref: $SRC not present (GPU box?) -- keeping existing $DST" >&2
  exit 0
fi
mkdir -p "$DST"
for f in model_clip.py model_ot.py utils_image.py; do
  install -m 0644 "$SRC/$f" "$DST/$f"
done
( cd "$DST" && sha256sum model_clip.py model_ot.py utils_image.py > SHA256SUMS )
echo "make_ref: copied reference loss-head modules into $DST"
