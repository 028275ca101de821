# Bit-exactness at BASELINE sizes: CPU HM vs GPUME=1 on 1080p / 4K clips, several cfgs and QPs, all pairs in parallel
# (16 host cores; the GPUME processes share the GPU under CUDA MPS -- without it they time-slice, see r1l_mps_sharing.log).
# usage: parity_scale.sh [all|ra]
cd /root/repo
export CUDA_MPS_PIPE_DIRECTORY=/tmp/mps_pipe CUDA_MPS_LOG_DIRECTORY=/tmp/mps_log
mkdir -p $CUDA_MPS_PIPE_DIRECTORY $CUDA_MPS_LOG_DIRECTORY
nvidia-cuda-mps-control -d
trap 'echo quit | nvidia-cuda-mps-control' EXIT
WHAT=${1:-all}
run() { # name cfg size frames qp bitdepth extra...
  name=$1; cfg=$2; size=$3; fr=$4; qp=$5; bd=$6; shift 6
  python tests/encode_compare.py --cfg $cfg --size $size --frames $fr --qp $qp --gpume 1 --bit-depth $bd -- "$@" > /tmp/ps_$name.json 2> /tmp/ps_$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open("/tmp/ps_%s.json" % n))
    print("%-28s bitstream_identical=%s recon_identical=%s cpu %.1f s gpu %.1f s  %s" % (n, d["bitstream_identical"], d["recon_identical"], d["cpu"]["wall_s"], d["gpu"]["wall_s"], d["gpu"]["gpume"][0][8:90]))
except Exception as e:
    print("%-28s FAILED %s %s" % (n, e, open("/tmp/ps_%s.err" % n).read()[-300:]))
PY
}
if [ "$WHAT" = all ]; then
run ldP_1080p_qp27 lowdelay_P_main 1920x1080 3 27 8 &
run ldP_1080p_qp37 lowdelay_P_main 1920x1080 3 37 8 &
run ldB_1080p_qp32 lowdelay_main 1920x1080 3 32 8 &
run ldP_1080p_fs2 lowdelay_P_main 1920x1080 2 32 8 --FastSearch=2 &
fi
run ra_1080p_qp32 randomaccess_main 1920x1080 9 32 8 --DecodingRefreshType=2 --IntraPeriod=16 &
run ra10_4k_qp32 randomaccess_main10 3840x2160 2 32 10 --DecodingRefreshType=2 --IntraPeriod=16 --SearchRange=128 &
wait
