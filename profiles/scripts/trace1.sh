cd /root/repo
python -c "
import sys; sys.path.insert(0,'hm-16.2_b200'); import synth; synth.write_yuv('/tmp/in.yuv',832,480,4,8)"
CFG=oracle/_ref/cfg/encoder_lowdelay_P_main.cfg
( time HMGPU_TRACE=1 hm-16.2_b200/host/build/TAppEncoderGpu -c $CFG -i /tmp/in.yuv -wdt 832 -hgt 480 -fr 30 -f 4 -q 32 -b /tmp/g.bin -o /tmp/g.yuv --GPUME=1 > /tmp/g.log ) 2>&1 | grep -E "real|GPUME|trace"
( time hm-16.2_b200/host/build/TAppEncoderGpu -c $CFG -i /tmp/in.yuv -wdt 832 -hgt 480 -fr 30 -f 4 -q 32 -b /tmp/g.bin -o /tmp/g.yuv --GPUME=1 > /tmp/g.log ) 2>&1 | grep -E "real|GPUME|trace"
( time oracle/_ref/TAppEncoderRef -c $CFG -i /tmp/in.yuv -wdt 832 -hgt 480 -fr 30 -f 4 -q 32 -b /tmp/c.bin -o /tmp/c.yuv > /tmp/c.log ) 2>&1 | grep real
md5sum /tmp/c.bin /tmp/g.bin
