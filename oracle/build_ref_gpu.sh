#!/bin/bash
# TEST INFRASTRUCTURE.  oracle/_ref/EAR_ref_gpu = the reference CLI with its Scene::Render thread fan-out
# (src/EAR.cpp:196-207) replaced by the binding of INTEGRATION.md (oracle/ref_gpu_stub.cpp) -> libear_b200.so.
# Everything else is the reference's own object code from oracle/build_ref.sh.  EAR.cpp is streamed through sed
# (nothing is copied into the repo): the opening line of the fan-out block becomes a call of the stub followed by
# `if (false)`, so the thread loop stays in the text and is never entered.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${EAR_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
  echo "build_ref_gpu.sh: no reference at $REF (expected on the GPU box) - keeping prebuilt $OUT" >&2
  exit 0
fi
bash "$HERE/build_ref.sh" > /dev/null
CXX="${CXX:-g++}"
FLAGS="-O2 -std=gnu++98 -fno-lifetime-dse -w -include math.h -I$HERE/shim -I$REF/src -I$REF/lib/wave -I$REF/lib/equalizer"
FIX='s/\b(animation|right_ear_animation|mesh|mat) > 0/\1 != 0/g'
HOOK='s/^\t\{std::vector<SceneContext>::const_iterator it = scs\.begin\(\);/\tRenderContextsOnGpu(scene, scs); if (false) {std::vector<SceneContext>::const_iterator it = scs.begin();/'
grep -q $'^\t{std::vector<SceneContext>::const_iterator it = scs.begin();' "$REF/src/EAR.cpp" || { echo "build_ref_gpu.sh: fan-out block not found in EAR.cpp" >&2; exit 1; }
{ echo 'class Scene; class SceneContext; void RenderContextsOnGpu(Scene*, std::vector<SceneContext>&);' ; sed -E -e "$FIX" -e "$HOOK" "$REF/src/EAR.cpp"; } \
  | $CXX $FLAGS -include vector -x c++ -c - -o "$OUT/obj/EAR_gpu.o"
$CXX $FLAGS -Dprivate=public -Dprotected=public -I"$ROOT/include" -c "$HERE/ref_gpu_stub.cpp" -o "$OUT/obj/ref_gpu_stub.o"
OBJS=""
for o in "$OUT"/obj/*.o; do
  case "$(basename "$o")" in EAR.o|EAR_gpu.o|ref_harness.o|ref_gpu_stub.o|ref_time_seed.o) ;; *) OBJS="$OBJS $o";; esac
done
$CXX -O2 -o "$OUT/EAR_ref_gpu" "$OUT/obj/EAR_gpu.o" "$OUT/obj/ref_gpu_stub.o" $OBJS "$OUT/obj/ref_time_seed.o" \
  -L"$ROOT/ear_b200/csrc" -lear_b200 -Wl,-rpath,'$ORIGIN/../../ear_b200/csrc' -lpthread
echo "built $OUT/EAR_ref_gpu"
