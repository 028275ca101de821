cd /root/repo
timeout 900 python -m pytest tests/test_encoder_md5.py -x -q 2>&1 | tail -3
python -c "
import sys; sys.path.insert(0,'hm-16.2_b200'); import synth; synth.write_yuv('/tmp/in.yuv',832,480,6,8)"
CFG=oracle/_ref/cfg/encoder_lowdelay_P_main.cfg
( time timeout 300 hm-16.2_b200/host/build/TAppEncoderGpu -c $CFG -i /tmp/in.yuv -wdt 832 -hgt 480 -fr 30 -f 6 -q 32 -b /tmp/g.bin -o /tmp/g.yuv --GPUME=1 > /tmp/g.log ) 2>&1 | grep -E "real|GPUME" | sed -E "s/candidates.*in hmgpu_me_search/.. in hmgpu_me_search/"
( time oracle/_ref/TAppEncoderRef -c $CFG -i /tmp/in.yuv -wdt 832 -hgt 480 -fr 30 -f 6 -q 32 -b /tmp/c.bin -o /tmp/c.yuv > /tmp/c.log ) 2>&1 | grep real
md5sum /tmp/c.bin /tmp/g.bin
grep -E "^POC" /tmp/c.log | sed -E 's/.*\[ET *([0-9]+) *\].*/cpu ET \1/' | tr '\n' ' '; echo
grep -E "^POC" /tmp/g.log | sed -E 's/.*\[ET *([0-9]+) *\].*/gpu ET \1/' | tr '\n' ' '; echo
