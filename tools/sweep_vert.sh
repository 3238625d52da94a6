# build variants of the vertical-counter kernel on the GPU box and time them
for v in "15 4" "15 5" "15 7" "12 7"; do
  set -- $v
  touch baseband_b200/csrc/bb_counts.cu
  NVCC_EXTRA="-DBB_VERT_WORDS=$1 -DBB_VERT_MINB=$2" python -m baseband_b200.build > /dev/null 2>&1
  echo "== words $1 minblocks $2"
  timeout 200 python tools/sweep_counts.py 1.0 2>&1 | grep "^counts [12] bit" 
done
touch baseband_b200/csrc/bb_counts.cu
python -m baseband_b200.build > /dev/null 2>&1
python -m pytest tests/test_gpu_counts.py -m gpu -q -x 2>&1 | tail -2
