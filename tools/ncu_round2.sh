set -x
NCU="ncu --set full --clock-control none --import-source on -f"
REPS=1 $NCU -k regex:k_decode_bitfield -o gpurun_out/r2_c2_base python tools/run_case.py c2 0.5 > gpurun_out/r2_ncu_b.log 2>&1
REPS=1 BB_TUNE_KNOCK=1 $NCU -k regex:k_decode_bitfield -o gpurun_out/r2_c2_knock1 python tools/run_case.py c2 0.5 >> gpurun_out/r2_ncu_b.log 2>&1
REPS=1 BB_TUNE_C2=1 $NCU -k regex:k_decode_bitfield -o gpurun_out/r2_c2_sel python tools/run_case.py c2 0.5 >> gpurun_out/r2_ncu_b.log 2>&1
REPS=1 BB_TUNE_C2=2 $NCU -k regex:k_decode_tile -o gpurun_out/r2_c2_tile python tools/run_case.py c2 0.5 >> gpurun_out/r2_ncu_b.log 2>&1
REPS=1 $NCU -k regex:k_mark4 -o gpurun_out/r2_c3_mark4 python tools/run_case.py mark4 0.5 >> gpurun_out/r2_ncu_b.log 2>&1
REPS=1 $NCU -k regex:k_int8 -o gpurun_out/r2_c4_guppi python tools/run_case.py guppi 0.5 >> gpurun_out/r2_ncu_b.log 2>&1
tail -3 gpurun_out/r2_ncu_b.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
