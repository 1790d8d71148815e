#!/bin/bash
# Sanitizer runs on a GPU box.  Writes logs under gpurun_out/sanitize/; summaries are copied to profiles/.
#   1. host ThreadSanitizer + AddressSanitizer builds of the library (host code only: -Xcompiler -fsanitize=...) under the native
#      stress driver tests/native/queue_stress.cpp (queue + lanes, 8 threads)
#   2. compute-sanitizer memcheck / racecheck / initcheck over the stress driver and a subset of the GPU parity tests
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
if [ "${1:-all}" = all ] || [ "$1" = host ]; then
  build_variant tsan -fsanitize=thread > $OUT/build_tsan.log 2>&1
  TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 history_size=4" timeout 600 /tmp/bpp_tsan/queue_stress 8 12 8 > $OUT/tsan.log 2>&1; echo "tsan rc=$?" | tee -a $OUT/summary.txt
  grep -c "WARNING: ThreadSanitizer" $OUT/tsan.log | sed 's/^/tsan warnings: /' | tee -a $OUT/summary.txt
  build_variant asan -fsanitize=address > $OUT/build_asan.log 2>&1
  ASAN_OPTIONS="protect_shadow_gap=0:detect_leaks=0:abort_on_error=0" timeout 600 /tmp/bpp_asan/queue_stress 8 12 8 > $OUT/asan.log 2>&1; echo "asan rc=$?" | tee -a $OUT/summary.txt
  grep -c "ERROR: AddressSanitizer" $OUT/asan.log | sed 's/^/asan errors: /' | tee -a $OUT/summary.txt
fi
if [ "${1:-all}" = all ] || [ "$1" = device ]; then
  g++ -O2 -std=c++17 tests/native/queue_stress.cpp -o /tmp/queue_stress -Lbulletproofs-plus_b200 -lbpp_b200 -Wl,-rpath,$PWD/bulletproofs-plus_b200 -lpthread
  for tool in memcheck racecheck initcheck; do
    timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 /tmp/queue_stress 2 3 2 > $OUT/cs_${tool}_stress.log 2>&1; echo "compute-sanitizer $tool stress rc=$?" | tee -a $OUT/summary.txt
    tail -3 $OUT/cs_${tool}_stress.log | tee -a $OUT/summary.txt
  done
  timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_queue.py -m gpu -x -q -k "create_multi or challenge or oversized" > $OUT/cs_memcheck_pytest.log 2>&1
  echo "compute-sanitizer memcheck pytest rc=$?" | tee -a $OUT/summary.txt; tail -4 $OUT/cs_memcheck_pytest.log | tee -a $OUT/summary.txt
fi
