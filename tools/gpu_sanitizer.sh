# compute-sanitizer over the GPU parity tests (results under
# gpurun_out/r2_sanitizer_*.txt); test_gpu_large is left out (64 GiB cases)
CS="compute-sanitizer --error-exitcode 77 --print-limit 20"
ALL="tests/test_gpu_bitfield.py tests/test_gpu_counts.py tests/test_gpu_index.py tests/test_gpu_corrupt.py tests/test_gpu_formats.py tests/test_gpu_streams.py tests/test_gpu_items.py"
SMEM="tests/test_gpu_bitfield.py tests/test_gpu_counts.py tests/test_gpu_index.py tests/test_gpu_formats.py"
run() {   # tool, limit, tests
    local out=gpurun_out/r2_sanitizer_$1.txt
    ( time timeout $2 $CS --tool $1 python -m pytest $3 -m gpu -q -x -p no:cacheprovider ) > $out 2>&1
    echo "$1 rc=$?" >> $out
    grep -E "passed|failed|SUMMARY|rc=|real" $out
}
run memcheck 1500 "$ALL"
run racecheck 1200 "$SMEM"
run synccheck 900 "$SMEM"
run initcheck 1200 "$SMEM"
