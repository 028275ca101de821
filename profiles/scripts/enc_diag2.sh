cd /root/repo
python -m pytest tests/test_encoder_md5.py -x -q 2>&1 | tail -3
python -c "
import sys; sys.path.insert(0,'hm-16.2_b200'); import synth; synth.write_yuv('/tmp/in1080.yuv',1920,1080,6,8)"
CFG=oracle/_ref/cfg/encoder_lowdelay_P_main.cfg
( time oracle/_ref/TAppEncoderRef -c $CFG -i /tmp/in1080.yuv -wdt 1920 -hgt 1080 -fr 30 -f 6 -q 32 -b /tmp/c.bin -o /tmp/c.yuv > /tmp/c.log ) 2>&1 | grep real &
( time hm-16.2_b200/host/build/TAppEncoderGpu -c $CFG -i /tmp/in1080.yuv -wdt 1920 -hgt 1080 -fr 30 -f 6 -q 32 -b /tmp/g.bin -o /tmp/g.yuv --GPUME=1 > /tmp/g.log ) 2>&1 | grep -E "real|GPUME"
wait
grep -E "^POC" /tmp/c.log | sed -E 's/.*\[ET *([0-9]+) *\].*/cpu ET \1/' | tr '\n' ' '; echo
grep -E "^POC" /tmp/g.log | sed -E 's/.*\[ET *([0-9]+) *\].*/gpu ET \1/' | tr '\n' ' '; echo
md5sum /tmp/c.bin /tmp/g.bin
python tests/encode_compare.py --cfg lowdelay_P_main --size 416x240 --frames 3 --gpume 1 -- --FastSearch=0 --SearchRange=64 | grep -E "wall_s|gpume|identical|GPUME"
