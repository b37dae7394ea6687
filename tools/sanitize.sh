#!/bin/bash
# Run ON THE GPU BOX (via gpurun): compute-sanitizer memcheck / racecheck / synccheck / initcheck over a 640x480
# extraction (TMA pyramid, extrema ring, descriptors), a 1024 x 1024 tensor-core match, a RANSAC homography and the
# device ImproveHomography.  Writes gpurun_out/sanitizer_<tag>.log.
#   gpurun -- 'bash tools/sanitize.sh r02'
tag=${1:-r02}
mkdir -p gpurun_out
log=gpurun_out/sanitizer_$tag.log
: > $log
cat > /tmp/san_run.py <<'PY'
import sys
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import cusift_b200 as csb
import parity_utils as PU
g1, g2 = PU.golden_frames()
ctx = csb.Context(0, 2)
p = csb.make_params(6, 0.0, 0.1)
k1 = ctx.extract(g1, p, max_pts=32768)
k2 = ctx.extract(g2, p, max_pts=32768, from_host=True)
ctx.extract(csb.synth(517, 389, 91), csb.make_params(4, 0.5, 0.3), max_pts=8192)          # odd size, partial tiles
r = np.random.default_rng(1)
def rs(n, seed):
    s = np.zeros(n, csb.SIFT_DTYPE)
    d = np.abs(np.random.default_rng(seed).standard_normal((n, 128))).astype(np.float32)
    s['data'] = d / np.linalg.norm(d, axis=1, keepdims=True)
    return s
m = ctx.match(rs(1024, 1), rs(1024, 2), 'l2')
m2 = ctx.match(k1[:3000], k2[:3000], 'l2')
rp = np.stack([r.choice(3000, 4, replace=False) for _ in range(256)], 1).astype(np.int32)
H, cnt = ctx.find_homography(m2, rp, 5.0)
H2, nf, _ = ctx.improve_homography(m2, H, 3, 0.0, 0.8, 3.0)
print('sanitizer workload ok', len(k1), len(k2), cnt, nf)
ctx.close()
PY
for tool in memcheck racecheck synccheck initcheck; do
  echo "===== compute-sanitizer --tool $tool" >> $log
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_run.py >> $log 2>&1
  echo "exit code: $?" >> $log
done
grep -E "=====|ERROR SUMMARY|RACECHECK SUMMARY|exit code|sanitizer workload" $log
