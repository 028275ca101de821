set -e
cd /root/repo
python -c "
import sys; sys.path.insert(0,'hm-16.2_b200'); import synth; synth.write_yuv('/tmp/in1080.yuv',1920,1080,2,8)"
CFG=oracle/_ref/cfg/encoder_lowdelay_P_main.cfg
( time oracle/_ref/TAppEncoderRef -c $CFG -i /tmp/in1080.yuv -wdt 1920 -hgt 1080 -fr 30 -f 2 -q 32 -b /tmp/c.bin -o /tmp/c.yuv > /tmp/c.log ) 2>&1 | grep real
grep -E "^POC|Total Time" /tmp/c.log | cut -c1-100
( time hm-16.2_b200/host/build/TAppEncoderGpu -c $CFG -i /tmp/in1080.yuv -wdt 1920 -hgt 1080 -fr 30 -f 2 -q 32 -b /tmp/g.bin -o /tmp/g.yuv --GPUME=1 > /tmp/g.log ) 2>&1 | grep -E "real|GPUME"
grep -E "^POC|Total Time" /tmp/g.log | cut -c1-100
