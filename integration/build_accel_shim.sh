#!/usr/bin/env bash
# Builds the Scene::accel_* shim (integration/scene_b2pt.inl): recompiles the reference's src/librender/scene.cpp — in the
# SCRATCH copy of the tree that oracle/build_ref.sh configured, with the two-line include patch quoted in the .inl and the
# compile / link commands ninja recorded for it — and installs the resulting libmitsuba-render.so under
# oracle/_ref/dist_b2accel/ (a second library directory put first on LD_LIBRARY_PATH by the test; oracle/_ref/dist stays the
# unmodified reference). Only possible where the configured scratch tree exists; the GPU box uses the prebuilt file.
set -euo pipefail
HERE=$(cd "$(dirname "$0")" && pwd); ROOT=$(dirname "$HERE")
SCRATCH=${SCRATCH:-/tmp/b2_refbuild}; SRC=$SCRATCH/src; B=$SCRATCH/build
DIST=$ROOT/oracle/_ref/dist; OUT=$ROOT/oracle/_ref/dist_b2accel
[ -f "$B/build.ninja" ] || { echo "build_accel_shim: no configured reference tree at $SCRATCH (run oracle/build_ref.sh)"; exit 0; }
[ -f "$DIST/libb2pt.so" ] || { echo "build_accel_shim: run integration/build_plugin.sh first"; exit 1; }
W=$(mktemp -d); trap 'rm -rf "$W"' EXIT
mkdir -p "$W/librender"
cp "$SRC"/src/librender/*.inl "$W/librender/"; cp "$HERE/scene_b2pt.inl" "$W/librender/"
python3 - "$SRC/src/librender/scene.cpp" "$W/librender/scene.cpp" <<'PY'
import sys
s = open(sys.argv[1]).read()
old = '#else\n#  include "scene_native.inl"'
assert old in s, "scene.cpp: include block not found"
open(sys.argv[2], "w").write(s.replace(old, '#elif defined(MTS_ENABLE_B2PT)\n#  include "scene_b2pt.inl"\n' + old, 1))
PY
cd "$B"
CC=$(ninja -t commands src/librender/CMakeFiles/mitsuba-render-obj.dir/scene.cpp.o | tail -1)
CC=${CC%% -MD*}                                         # compiler + flags, without the dependency / output part
$CC -DMTS_ENABLE_B2PT -I"$ROOT/include" -o "$W/scene.cpp.o" -c "$W/librender/scene.cpp"
LINK=$(ninja -t commands src/librender/libmitsuba-render.so | tail -1)
LINK=${LINK#: && }; LINK=${LINK%% && :*}
LINK=${LINK//src\/librender\/CMakeFiles\/mitsuba-render-obj.dir\/scene.cpp.o/$W/scene.cpp.o}
LINK=${LINK/-o src\/librender\/libmitsuba-render.so/-o $W/libmitsuba-render.so}
LINK=${LINK/-Wl,--dependency-file=src\/librender\/CMakeFiles\/mitsuba-render.dir\/link.d/}
eval "$LINK -L$DIST -lb2pt"
mkdir -p "$OUT"; cp "$W/libmitsuba-render.so" "$OUT/libmitsuba-render.so"
echo "build_accel_shim: installed $OUT/libmitsuba-render.so"
