set -x
nproc; lscpu | grep -E "Model name|MHz|L3|NUMA node\(s\)"; which perf gprof valgrind; free -g | head -2
cd /root/repo
python profiles/latency_probe.py 2>&1 | tee gpurun_out/latency_probe_r1d.txt
python -c "
import sys; sys.path.insert(0,'hm-16.2_b200'); import synth; synth.write_yuv('/tmp/in1080.yuv',1920,1080,3,8)"
CFG=oracle/_ref/cfg/encoder_lowdelay_P_main.cfg
( time oracle/_ref/TAppEncoderRef -c $CFG -i /tmp/in1080.yuv -wdt 1920 -hgt 1080 -fr 30 -f 3 -q 32 -b /tmp/c.bin -o /tmp/c.yuv > /tmp/c.log ) 2>&1 | grep real
grep -E "^POC|Total Time" /tmp/c.log | cut -c1-110
( time hm-16.2_b200/host/build/TAppEncoderGpu -c $CFG -i /tmp/in1080.yuv -wdt 1920 -hgt 1080 -fr 30 -f 3 -q 32 -b /tmp/g.bin -o /tmp/g.yuv --GPUME=1 > /tmp/g.log ) 2>&1 | grep -E "real|GPUME"
grep -E "^POC|Total Time" /tmp/g.log | cut -c1-110
( time hm-16.2_b200/host/build/TAppEncoderGpu -c $CFG -i /tmp/in1080.yuv -wdt 1920 -hgt 1080 -fr 30 -f 3 -q 32 -b /tmp/g0.bin -o /tmp/g0.yuv --GPUME=0 > /tmp/g0.log ) 2>&1 | grep -E "real|GPUME"
grep -E "^POC|Total Time" /tmp/g0.log | cut -c1-110
md5sum /tmp/c.bin /tmp/g.bin /tmp/g0.bin
