# round 2: evidence gaps of VERDICT r1 item 6: BASELINE-size MD5 tests, ncu --set full of the kernels outside the ME step,
# compute-sanitizer racecheck / memcheck of the mailbox server and of the CTU-group kernel
cd /root/repo
S=$(date +%s)
timeout 1500 python -m pytest tests/test_encoder_md5.py -m gpu -q -k "1080p or 4k" --durations=5 2>&1 | tail -12 > gpurun_out/r2l_pytest_slow.log; tail -5 gpurun_out/r2l_pytest_slow.log; echo "slow tests $(( $(date +%s) - S )) s"
timeout 300 python profiles/prof_misc.py 3 > gpurun_out/r2l_prof_misc.txt 2>&1; cat gpurun_out/r2l_prof_misc.txt | tail -12
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"phase_planes|predict_kernel|pred_error|fwd_transform|quant_kernel|dist_batch|pad_convert" -c 14 -o gpurun_out/r2l_misc python profiles/prof_misc.py 1 > gpurun_out/r2l_ncu_misc.log 2>&1; tail -2 gpurun_out/r2l_ncu_misc.log | cut -c1-160
S=$(date +%s)
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mailbox_server or fractional_window_kernel" > gpurun_out/r2l_racecheck.log 2>&1; tail -6 gpurun_out/r2l_racecheck.log; echo "racecheck $(( $(date +%s) - S )) s"
S=$(date +%s)
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mailbox_server or fractional_window_kernel or fractional_kernels_agree" > gpurun_out/r2l_memcheck.log 2>&1; tail -6 gpurun_out/r2l_memcheck.log; echo "memcheck $(( $(date +%s) - S )) s"
