cd /root/repo
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "low_latency or errors" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_encoder_md5.py -x -q 2>&1 | tail -3
python -c "
import sys; sys.path.insert(0,'hm-16.2_b200'); import synth; synth.write_yuv('/tmp/in.yuv',832,480,4,8)"
CFG=oracle/_ref/cfg/encoder_lowdelay_P_main.cfg
for srv in 1 0; do
( time HMGPU_SERVER=$srv HMGPU_SERVER_STATS=1 timeout 300 hm-16.2_b200/host/build/TAppEncoderGpu -c $CFG -i /tmp/in.yuv -wdt 832 -hgt 480 -fr 30 -f 4 -q 32 -b /tmp/g$srv.bin -o /tmp/g$srv.yuv --GPUME=1 > /tmp/g.log ) 2>&1 | grep -E "real|GPUME|server"
done
( time oracle/_ref/TAppEncoderRef -c $CFG -i /tmp/in.yuv -wdt 832 -hgt 480 -fr 30 -f 4 -q 32 -b /tmp/c.bin -o /tmp/c.yuv > /tmp/c.log ) 2>&1 | grep real
md5sum /tmp/c.bin /tmp/g1.bin /tmp/g0.bin
timeout 200 python profiles/latency_probe.py 2>&1 | head -8
