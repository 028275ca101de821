cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fractional_kernels_agree or frac_only or tz_search_and_frac" 2>&1 | tail -25 > gpurun_out/r2d_pytest.log; tail -5 gpurun_out/r2d_pytest.log
for tma in 0 1; do echo "HMGPU_FRAC_TMA=$tma"; HMGPU_FRAC_TMA=$tma timeout 300 python profiles/frac_order_probe.py 2>&1 | tail -4; done
for swz in 1 0; do for v in 0 1 3; do echo "swz $swz variant $v"; HMGPU_FRAC3_SWIZZLE=$swz HMGPU_FRAC3_VARIANT=$v timeout 300 python profiles/frac_order_probe.py 2>&1 | tail -3 | head -2; done; done
timeout 300 python profiles/prof_misc.py 3 2>&1 | tail -12
