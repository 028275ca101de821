cd /root/repo
which nvidia-cuda-mps-control nvidia-cuda-mps-server
export CUDA_MPS_PIPE_DIRECTORY=/tmp/mps_pipe CUDA_MPS_LOG_DIRECTORY=/tmp/mps_log
mkdir -p $CUDA_MPS_PIPE_DIRECTORY $CUDA_MPS_LOG_DIRECTORY
nvidia-cuda-mps-control -d; echo "mps start rc=$?"
sleep 1
python -c "
import sys; sys.path.insert(0,'hm-16.2_b200'); import synth; synth.write_yuv('/tmp/in.yuv',832,480,3,8)"
CFG=oracle/_ref/cfg/encoder_lowdelay_P_main.cfg
one() { ( time timeout 600 hm-16.2_b200/host/build/TAppEncoderGpu -c $CFG -i /tmp/in.yuv -wdt 832 -hgt 480 -fr 30 -f 3 -q 32 -b /tmp/g$1.bin -o /tmp/g$1.yuv --GPUME=1 > /tmp/g$1.log ) 2>&1 | grep -E "real" | sed "s/^/proc $1 /"; }
for mode in 1 0; do
  export HMGPU_SERVER=$mode
  echo "== 1 process, HMGPU_SERVER=$mode (under MPS)"; one a
  echo "== 4 processes sharing the GPU, HMGPU_SERVER=$mode (under MPS)"
  one a & one b & one c & one d & wait
done
md5sum /tmp/ga.bin /tmp/gb.bin /tmp/gc.bin /tmp/gd.bin | awk '{print $1}' | sort -u | wc -l
echo quit | nvidia-cuda-mps-control; echo "mps quit rc=$?"
tail -5 /tmp/mps_log/control.log 2>/dev/null
echo "== without MPS: 4 processes, launch path"
export HMGPU_SERVER=0
one a & one b & one c & one d & wait
