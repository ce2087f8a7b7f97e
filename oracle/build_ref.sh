#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the UNMODIFIED reference (mitsuba2 @ /root/reference) for the
# scalar_rgb variant into oracle/_ref/ (git-ignored; travels to the GPU box with gpurun).
# Nothing from the reference's sources is copied into the repo: the source tree is copied to
# a scratch dir under /tmp (the configure step writes mitsuba.conf / config.h into the source
# dir, and /root/reference is read-only), built there, and only BINARIES (+ the scene data
# files needed to run the CLI) are installed under oracle/_ref/.
#
# Recipe = SURVEY.md Appendix B (three build-only tweaks, no source change on the hot path):
#   1. mitsuba.conf enables only "scalar_rgb"
#   2. ext/CMakeLists.txt: use the vendored libpng/libjpeg on Linux too (no system libpng-dev)
#   3. -fpermissive -Wno-changes-meaning for vendored TBB 2019 under GCC 13; python/gui off
# Used by: tests/golden/make_golden.py (fixture generation, this container only) and
#          bench.py --impl reference (cpu_baseline.kind == "reference").
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
SCRATCH=${SCRATCH:-/tmp/b2_refbuild}
JOBS=${JOBS:-$(nproc)}
if [ ! -d "$REF" ]; then echo "build_ref: $REF absent (GPU box?) — using prebuilt $OUT if present"; exit 0; fi
if [ -x "$OUT/dist/mitsuba" ] && [ -z "${FORCE:-}" ]; then echo "build_ref: $OUT/dist/mitsuba exists"; exit 0; fi
mkdir -p "$SCRATCH"
if [ ! -d "$SCRATCH/src" ]; then cp -r "$REF" "$SCRATCH/src"; chmod -R u+w "$SCRATCH/src"; fi
cd "$SCRATCH/src"
python3 - <<'PY'
import re
t = open('resources/mitsuba.conf.template').read()
# keep only scalar_rgb in "enabled" and make it the default
t = re.sub(r'"enabled": \[.*?\]', '"enabled": [\n        "scalar_rgb"\n    ]', t, flags=re.S)
t = re.sub(r'"default": "[a-z_]+"', '"default": "scalar_rgb"', t)
open('mitsuba.conf', 'w').write(t)
c = open('ext/CMakeLists.txt').read()
c = c.replace('if (WIN32 OR APPLE)', 'if (TRUE)')
open('ext/CMakeLists.txt', 'w').write(c)
PY
B=$SCRATCH/build
cmake -GNinja -S "$SCRATCH/src" -B "$B" \
  -DCMAKE_POLICY_VERSION_MINIMUM=3.5 -DCMAKE_BUILD_TYPE=Release \
  -DMTS_ENABLE_PYTHON=OFF -DMTS_ENABLE_GUI=OFF -DMTS_ENABLE_EMBREE=OFF \
  "-DCMAKE_CXX_FLAGS=-fpermissive -Wno-changes-meaning" \
  "-DCMAKE_EXE_LINKER_FLAGS=-Wl,-rpath-link,$B/ext_build/libpng:$B/ext_build/libjpeg" \
  > "$SCRATCH/configure.log" 2>&1
ninja -C "$B" -j "$JOBS" > "$SCRATCH/build.log" 2>&1
rm -rf "$OUT"; mkdir -p "$OUT"
cp -r "$B/dist" "$OUT/dist"
# vendored png/jpeg shared libs live outside dist/
find "$B/ext_build/libpng" "$B/ext_build/libjpeg" -name '*.so*' -exec cp -P {} "$OUT/dist/" \;
# scene DATA needed to drive the CLI on the GPU box (not source code; stays git-ignored)
mkdir -p "$OUT/scenes"
cp -r "$REF/resources/data/scenes/cbox" "$OUT/scenes/cbox"
cp -r "$REF/resources/data/scenes/matpreview" "$OUT/scenes/matpreview"     # shipped scene of the drop-in test (BASELINE configs[2])
echo "g++ $(g++ -dumpversion); flags: reference defaults (-O3 -march=native -ffp-contract=fast); host: $(grep -m1 'model name' /proc/cpuinfo | cut -d: -f2)" > "$OUT/BUILD_INFO.txt"
# the reference's CMake adds -march=native: record the ISA extensions of THIS host so that a box lacking one of them can
# say so loudly (tests/ref_cli.py: isa_missing(); a SIGILL of the CLI is reported with this list) instead of failing obscurely
grep -m1 '^flags' /proc/cpuinfo | tr ' ' '\n' | grep -E '^(avx|fma|bmi|sse4|f16c|popcnt|adx|sha_ni|vaes|vpclmul|gfni)' | sort -u | tr '\n' ' ' > "$OUT/ISA_FLAGS.txt"
echo "build_ref: installed $OUT/dist"
