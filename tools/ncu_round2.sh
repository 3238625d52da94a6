# ncu --set full captures of the final kernels (one launch each)
set -x
NCU="ncu --set full --clock-control none --import-source on -f"
REPS=1 $NCU -k regex:k_decode_bitfield -o gpurun_out/r2_final_c2_decode python tools/run_case.py c2 0.5 > gpurun_out/r2_ncu_final.log 2>&1
REPS=1 $NCU -k regex:k_encode_bitfield -o gpurun_out/r2_final_c2_encode python tools/run_case.py c2 0.5 >> gpurun_out/r2_ncu_final.log 2>&1
REPS=1 $NCU -k regex:k_mark4_decode -o gpurun_out/r2_final_c3_mark4 python tools/run_case.py mark4 0.5 >> gpurun_out/r2_ncu_final.log 2>&1
REPS=1 $NCU -k regex:k_int8_decode -o gpurun_out/r2_final_c4_guppi python tools/run_case.py guppi 0.5 >> gpurun_out/r2_ncu_final.log 2>&1
REPS=1 $NCU -k regex:k_decode_bitfield -o gpurun_out/r2_final_c5_mark5b python tools/run_case.py mark5b 0.5 >> gpurun_out/r2_ncu_final.log 2>&1
REPS=1 $NCU -k regex:k_state_counts -o gpurun_out/r2_final_counts python tools/run_case.py counts 0.5 >> gpurun_out/r2_ncu_final.log 2>&1
tail -3 gpurun_out/r2_ncu_final.log
