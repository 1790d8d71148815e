#!/bin/bash
# Sanitizer runs on a GPU box.  ONE tool per gpurun call (B200_PROFILING.md: several compute-sanitizer tools in
# one call have left GPUs unusable).  Logs go to gpurun_out/sanitize/; summaries are copied to profiles/ by hand.
#   scripts/sanitize.sh host        ThreadSanitizer, then AddressSanitizer: host code of the library rebuilt with -fsanitize=..., under
#                                   the native stress driver tests/native/queue_stress.cpp (queue + lanes, 8 threads each)
#   scripts/sanitize.sh memcheck|racecheck|initcheck      compute-sanitizer --tool <that> over the stress driver (small sizes)
set -u
cd ${GRAFT_REPO_ROOT:-$(dirname $0)/..}
OUT=gpurun_out/sanitize; mkdir -p $OUT
CSRC=bulletproofs-plus_b200/csrc
ARCH="-gencode arch=compute_100a,code=sm_100a"
build_variant() {   # $1 = name, $2 = sanitizer flag
  local B=/tmp/bpp_$1; mkdir -p $B
  for f in $CSRC/*.cu; do nvcc $ARCH -O1 -g -std=c++17 -lineinfo -Xcompiler -fPIC,-fno-omit-frame-pointer,$2 -c $f -o $B/$(basename $f).o & done
  for f in $CSRC/*.cpp; do g++ -O1 -g -std=c++17 -fPIC -fno-omit-frame-pointer $2 -c $f -o $B/$(basename $f).o & done
  wait
  g++ -shared -o $B/libbpp_b200.so $B/*.o $2 -L/usr/local/cuda/lib64 -lcudart -lpthread -ldl -lrt
  g++ -O1 -g -std=c++17 $2 -fno-omit-frame-pointer tests/native/queue_stress.cpp -o $B/queue_stress -L$B -lbpp_b200 -Wl,-rpath,$B -lpthread
}
case "${1:-host}" in
host)
  build_variant tsan -fsanitize=thread > $OUT/build_tsan.log 2>&1
  TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 history_size=4" timeout 900 /tmp/bpp_tsan/queue_stress 8 12 8 > $OUT/tsan.log 2>&1; echo "tsan rc=$?" | tee $OUT/summary_host.txt
  echo "tsan warnings: $(grep -c 'WARNING: ThreadSanitizer' $OUT/tsan.log)" | tee -a $OUT/summary_host.txt
  grep -A12 "WARNING: ThreadSanitizer" $OUT/tsan.log | grep "#0\|#1\|#2\|WARNING" | sort | uniq -c | sort -rn | head -30 | tee -a $OUT/summary_host.txt
  build_variant asan -fsanitize=address > $OUT/build_asan.log 2>&1
  ASAN_OPTIONS="protect_shadow_gap=0:detect_leaks=0:abort_on_error=0" timeout 900 /tmp/bpp_asan/queue_stress 8 12 8 > $OUT/asan.log 2>&1; echo "asan rc=$?" | tee -a $OUT/summary_host.txt
  echo "asan errors: $(grep -c 'ERROR: AddressSanitizer' $OUT/asan.log)" | tee -a $OUT/summary_host.txt
  tail -3 $OUT/tsan.log $OUT/asan.log | tee -a $OUT/summary_host.txt
  ;;
memcheck|racecheck|initcheck)
  tool=$1
  g++ -O2 -std=c++17 tests/native/queue_stress.cpp -o /tmp/queue_stress -Lbulletproofs-plus_b200 -lbpp_b200 -Wl,-rpath,$PWD/bulletproofs-plus_b200 -lpthread
  /tmp/queue_stress 2 3 2 > $OUT/plain_stress.log 2>&1 && \
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 /tmp/queue_stress 2 3 2 > $OUT/cs_${tool}_stress.log 2>&1
  echo "compute-sanitizer $tool over queue_stress 2 3 2: rc=$?" | tee $OUT/summary_$tool.txt
  tail -6 $OUT/cs_${tool}_stress.log | tee -a $OUT/summary_$tool.txt
  ;;
esac
