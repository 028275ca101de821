cd /root/repo
python -c "
import sys; sys.path.insert(0,'hm-16.2_b200'); import synth; synth.write_yuv('/tmp/in.yuv',832,480,4,8)"
CFG=oracle/_ref/cfg/encoder_lowdelay_P_main.cfg
for idle in 50 200 1000 5000; do
echo idle $idle
( time HMGPU_SERVER_IDLE_US=$idle HMGPU_SERVER_STATS=1 timeout 300 hm-16.2_b200/host/build/TAppEncoderGpu -c $CFG -i /tmp/in.yuv -wdt 832 -hgt 480 -fr 30 -f 4 -q 32 -b /tmp/g.bin -o /tmp/g.yuv --GPUME=1 > /tmp/g.log ) 2>&1 | grep -E "real|GPUME|server" | sed -E 's/candidates.*in hmgpu_me_search/.. in hmgpu_me_search/; s/one-time.*//'
done
