#!/usr/bin/env bash
# Builds integration/b2path.cpp (the reference-side binding of libb2pt) into
# oracle/_ref/dist/plugins/b2path.so, against the reference headers of the scratch tree that
# oracle/build_ref.sh configured (it holds the generated include/mitsuba/core/config.h), with the
# flags the reference uses for its own integrator plugins (ninja -t commands src/integrators/path.so).
# Only possible where /root/reference exists; the GPU box uses the prebuilt plugin.
set -euo pipefail
HERE=$(cd "$(dirname "$0")" && pwd); ROOT=$(dirname "$HERE")
SRC=${SCRATCH:-/tmp/b2_refbuild}/src
DIST=$ROOT/oracle/_ref/dist
[ -d "$SRC/include/mitsuba" ] && [ -f "$SRC/include/mitsuba/core/config.h" ] || { echo "build_plugin: no configured reference tree at $SRC (run oracle/build_ref.sh)"; exit 0; }
[ -f "$ROOT/mitsuba2_b200/libb2pt.so" ] || { echo "build_plugin: build libb2pt.so first"; exit 1; }
cp -f "$ROOT/mitsuba2_b200/libb2pt.so" "$DIST/libb2pt.so"          # next to libmitsuba-*.so: found through the $ORIGIN rpath
g++ -DLITTLE_ENDIAN -D__LINUX__ -I"$SRC/include" -I"$SRC/ext/tinyformat" -I"$SRC/ext/enoki/include" -I"$SRC/ext/tbb/include" \
    -I"$ROOT/include" -fpermissive -Wno-changes-meaning -std=gnu++17 -O3 -DNDEBUG -fPIC -fno-stack-protector \
    -fomit-frame-pointer -fno-math-errno -ffp-contract=fast -march=native -finline-limit=150 \
    -shared -Wl,-soname,b2path.so -o "$DIST/plugins/b2path.so" "$HERE/b2path.cpp" \
    -Wl,-rpath,'$ORIGIN/..' -L"$DIST" -lmitsuba-render -lmitsuba-core -lb2pt -ltbb -lpthread -ldl
# the same adapter as `<integrator type="b2direct">` (DirectIntegrator with shading_samples = 1 == path with max_depth 2)
TMPD=$(mktemp -d); sed 's/B2PathIntegrator/B2DirectIntegrator/g' "$HERE/b2path.cpp" > "$TMPD/b2direct.cpp"
g++ -DB2_DIRECT -DLITTLE_ENDIAN -D__LINUX__ -I"$SRC/include" -I"$SRC/ext/tinyformat" -I"$SRC/ext/enoki/include" -I"$SRC/ext/tbb/include" \
    -I"$ROOT/include" -fpermissive -Wno-changes-meaning -std=gnu++17 -O3 -DNDEBUG -fPIC -fno-stack-protector \
    -fomit-frame-pointer -fno-math-errno -ffp-contract=fast -march=native -finline-limit=150 \
    -shared -Wl,-soname,b2direct.so -o "$DIST/plugins/b2direct.so" "$TMPD/b2direct.cpp" \
    -Wl,-rpath,'$ORIGIN/..' -L"$DIST" -lmitsuba-render -lmitsuba-core -lb2pt -ltbb -lpthread -ldl
rm -rf "$TMPD"
echo "build_plugin: installed $DIST/plugins/b2path.so and b2direct.so"
