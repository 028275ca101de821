cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "transform or quant" 2>&1 | tail -3
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"quant_kernel|intra_costs|sao_" -c 8 -o gpurun_out/r2s_misc2 python profiles/prof_misc.py 1 > gpurun_out/r2s_ncu_misc2.log 2>&1; tail -1 gpurun_out/r2s_ncu_misc2.log | cut -c1-160
