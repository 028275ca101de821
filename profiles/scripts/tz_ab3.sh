cd /root/repo
HMGPU_TZ_CARVE=50 HMGPU_TZ_MERGE=1 HMGPU_TZ_SPLIT=5 python profiles/tz_ab.py 2>&1 | tail -2 | cut -c1-120
HMGPU_TZ_P2=0 HMGPU_TZ_CARVE=50 HMGPU_TZ_MERGE=1 HMGPU_TZ_SPLIT=5 python profiles/tz_ab.py 2>&1 | tail -2 | cut -c1-120
HMGPU_TZ_P2=0 HMGPU_TZ_MERGE=1 HMGPU_TZ_SPLIT=5 python profiles/tz_ab.py 2>&1 | tail -2 | cut -c1-120
