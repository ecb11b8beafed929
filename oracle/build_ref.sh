#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Compiles the UNMODIFIED reference (fpsample v1.0.2) from the sources
# where they lie under /root/reference into oracle/_ref/fpsample_ref/ (git-ignored, shipped to the
# GPU box by gpurun).  No reference source is copied into the repo: only the built extension and a
# byte-copy of the reference's pure-python front-end land in the ignored directory.
# Mirrors the reference's own flags: CMakeLists.txt:13-19 (C++17, no -march, no fast-math),
# scikit-build-core Release default -O3 -DNDEBUG.
set -euo pipefail
REF=${1:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
DST="$HERE/_ref/fpsample_ref"
if [ ! -f "$REF/src/lib.cpp" ]; then
  echo "build_ref: $REF/src/lib.cpp not found (expected on the GPU box); keeping prebuilt files" >&2
  exit 0
fi
mkdir -p "$DST"
PYINC=$(python3 -c "import sysconfig;print(sysconfig.get_paths()['include'])")
PBINC=$(python3 -c "import pybind11;print(pybind11.get_include())")
EXT=$(python3 -c "import sysconfig;print(sysconfig.get_config_var('EXT_SUFFIX'))")
OUT="$DST/_fpsample$EXT"
if [ ! -f "$OUT" ] || [ "$REF/src/lib.cpp" -nt "$OUT" ]; then
  g++ -O3 -DNDEBUG -std=c++17 -fPIC -shared -fvisibility=hidden -ffp-contract=off \
      -I"$PYINC" -I"$PBINC" -I"$REF/src" -DVERSION_INFO=1.0.2 \
      "$REF/src/lib.cpp" -o "$OUT"
fi
install -m 644 "$REF/src/fpsample/__init__.py" "$DST/__init__.py"
# Plain C-ABI build of the reference's wrapper.hpp (extern "C" bucket_fps_kdline), no python needed:
CABI="$HERE/_ref/libfpsample_ref_cabi.so"
if [ ! -f "$CABI" ]; then
  printf '#include "wrapper.hpp"\n' > "$HERE/_ref/_cabi_tu.cpp"
  g++ -O3 -DNDEBUG -std=c++17 -fPIC -shared -ffp-contract=off -I"$REF/src" \
      "$HERE/_ref/_cabi_tu.cpp" -o "$CABI"
  rm -f "$HERE/_ref/_cabi_tu.cpp"
fi
echo "build_ref: ok -> $DST"
