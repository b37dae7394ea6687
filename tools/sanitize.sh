#!/bin/bash
# Run ON THE GPU BOX (via gpurun): compute-sanitizer memcheck / racecheck / synccheck / initcheck over a 640x480
# extraction (TMA pyramid, extrema ring, descriptors), a 1024 x 1024 tensor-core match, a RANSAC homography, the
# device ImproveHomography (cluster kernel), the batched all-pairs path and the exact redo kernels.  Writes gpurun_out/sanitizer_<tag>.log.
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
# all-pairs path (k_ransac_prep, k_score with the fused arg-max, cluster ImproveHomography) on three sets, and a match
# whose short lists overflow (exact duplicates), so that the redo kernels run
sets = [np.ascontiguousarray(k1[:1536]), np.ascontiguousarray(k2[:1536]), np.ascontiguousarray(k1[1536:3072])]
dp = [ctx.upload_sift(s) for s in sets]
ap = ctx.allpairs(dp, [len(s) for s in sets], csb.all_pairs(3), 'l2', 256, 0.0, 0.80, 5.0, 7, improve_loops=3)
a, b = rs(600, 5), rs(2048, 6)
for c in range(40, 40 + 21 * 16, 16):
    b['data'][c] = a['data'][5]
before = csb.lib().csb_match_redo_blocks(ctx.h)
m3 = ctx.match(a, b, 'l2')
print('sanitizer workload ok', len(k1), len(k2), cnt, nf, 'redo blocks', csb.lib().csb_match_redo_blocks(ctx.h) - before)
ctx.close()
PY
for tool in memcheck racecheck synccheck initcheck; do
  echo "===== compute-sanitizer --tool $tool" >> $log
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_run.py >> $log 2>&1
  echo "exit code: $?" >> $log
done
grep -E "=====|ERROR SUMMARY|RACECHECK SUMMARY|exit code|sanitizer workload" $log
