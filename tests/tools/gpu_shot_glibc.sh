# Round-2 GPU validation of prefs.devices.b200.libm = 'glibc': the gpurun calls of the last session,
# in order (each was `gpurun -- 'bash tests/tools/gpu_shot_glibc.sh'` with that call's body; outputs
# merged into gpurun_out/ and copied to profiles/ under the names given in profiles/README.md R2.2).
#
# call 1 (104 s charged)  pytest tests/test_parity_gpu.py -k glibc_math                -> r02_glibc_math_gpu_tests.log
#                         python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs
#                         python bench.py --workload cobahh_256k --libm glibc --steps 5 --warmup 3 --no-cpu-baseline
#                                                                                      -> r02_bench_cobahh_256k_glibc.json
# call 2 (106 s)          python bench.py --workload cobahh_256k --libm glibc --phases --steps 3 --warmup 3 --no-cpu-baseline
#                                                                                      -> r02_phases_cobahh_256k_glibc.txt
#                         python bench.py --gpus 1 --steps 20 --warmup 5               -> r02_bench_default_final.json
# call 3 (79 s)           pytest tests/test_parity_gpu.py -k "glibc_math or cuba_1000-True or cobahh_1000-False
#                                 or test_device_math_identical"; smoke()              -> r02_glibc_math_gpu_tests2.log
# call 4 (47 s), call 5 (47 s): the body below (after tanh/sinh/cosh, then after sin/cos were added)
#                                                                                      -> r02_glibc_math_gpu_tests3.log, ...tests4.log
mkdir -p gpurun_out
( time timeout 80 python -m pytest tests/test_parity_gpu.py -q -x -k "every_libm_call or cuba_1000-True" ) > gpurun_out/r2v_tests.log 2>&1
tail -8 gpurun_out/r2v_tests.log
