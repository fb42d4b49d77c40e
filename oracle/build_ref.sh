#!/bin/bash
# TEST INFRASTRUCTURE.  Compiles the reference's own C++ sources, from where they lie
# under $EAR_REFERENCE (default /root/reference), into oracle/_ref/ (git-ignored):
#   oracle/_ref/EAR_ref       the reference CLI (`render`, `calc T60`, `test`)
#   oracle/_ref/ref_harness   oracle/ref_harness.cpp linked against the same objects
# No reference source is copied into the repo: each .cpp is streamed through sed
# straight into g++ (stdin).  The sed rewrites the 10 ordered pointer-vs-zero
# comparisons (`ptr > 0`, ill-formed since C++11/gcc 11; src/MonoRecorder.cpp:101,103,
# src/StereoRecorder.cpp:134,136,143, src/SoundFile.cpp:138,143,216,231,
# src/Scene.cpp:161) to `ptr != 0`; nothing else changes.
# Flags: -fno-lifetime-dse is REQUIRED (src/Scene.cpp:71-72 reads `refl` inside its
# own initialiser; modern gcc otherwise drops the store and every bounce is NaN);
# -include math.h supplies fabs/pow overloads the sources rely on transitively.
# GMTL 0.6.1 and boost::thread are not vendored by the reference: oracle/shim/ has
# stand-ins (see the headers there).  FFTW is left off (direct convolution).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${EAR_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
  echo "build_ref.sh: no reference at $REF (expected on the GPU box) - keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT/obj"
CXX="${CXX:-g++}"
FLAGS="-O2 -std=gnu++98 -fno-lifetime-dse -w -include math.h -I$HERE/shim -I$REF/src -I$REF/lib/wave -I$REF/lib/equalizer"
FIX='s/\b(animation|right_ear_animation|mesh|mat) > 0/\1 != 0/g'
OBJS=""
for f in "$REF"/src/*.cpp "$REF/lib/wave/WaveFile.cpp" "$REF/lib/equalizer/Equalizer.cpp"; do
  o="$OUT/obj/$(basename "${f%.cpp}").o"
  if [ "$(basename "$f")" != "EAR.cpp" ]; then OBJS="$OBJS $o"; fi
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ "$HERE/shim/gmtl/gmtl.h" -nt "$o" ] || [ "$HERE/shim/boost/thread/thread.hpp" -nt "$o" ]; then
    sed -E "$FIX" "$f" | $CXX $FLAGS -x c++ -c - -o "$o"
  fi
done
$CXX $FLAGS -c "$HERE/ref_time_seed.cpp" -o "$OUT/obj/ref_time_seed.o"
$CXX -O2 -o "$OUT/EAR_ref" "$OUT/obj/EAR.o" $OBJS "$OUT/obj/ref_time_seed.o" -lpthread
$CXX $FLAGS -c "$HERE/ref_harness.cpp" -o "$OUT/obj/ref_harness.o"
$CXX -O2 -o "$OUT/ref_harness" "$OUT/obj/ref_harness.o" $OBJS -lpthread
echo "built $OUT/EAR_ref and $OUT/ref_harness"
