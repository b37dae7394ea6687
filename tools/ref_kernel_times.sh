#!/bin/bash
# Run ON THE GPU BOX (via gpurun): ncu launch lists of the UNMODIFIED reference (oracle/_ref/ref_driver) on the bench
# frame and on an 8192 x 8192 match, so every product kernel has the reference's kernel time beside it.
#   gpurun -- 'bash tools/ref_kernel_times.sh r02'
tag=${1:-r02}
mkdir -p gpurun_out /tmp/refk
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
import cusift_b200 as csb
from oracle import oracle as O
csb.synth(1920, 1080, 1000).tofile('/tmp/refk/f1080.f32')
def rand_set(n, seed):
    r = np.random.default_rng(seed)
    s = np.zeros(n, csb.SIFT_DTYPE)
    d = np.abs(r.standard_normal((n, 128))).astype(np.float32)
    s['data'] = d / np.linalg.norm(d, axis=1, keepdims=True)
    return s
O.write_sift_file('/tmp/refk/a.sift', rand_set(8192, 1))
O.write_sift_file('/tmp/refk/b.sift', rand_set(8192, 2))
PY
NCU="ncu --clock-control none --metrics gpu__time_duration.sum --csv"
# 3 frames: the summariser keeps the last one (warm)
timeout 300 $NCU --log-file gpurun_out/ref_launches_extract_$tag.csv oracle/_ref/ref_driver bench /tmp/refk/f1080.f32 1920 1080 5 0.0 1.0 10.0 0.0 16384 2 1 > gpurun_out/ref_launches_extract_$tag.log 2>&1
timeout 300 $NCU --log-file gpurun_out/ref_launches_match_$tag.csv oracle/_ref/ref_driver benchmatch /tmp/refk/a.sift /tmp/refk/b.sift 1 1 > gpurun_out/ref_launches_match_$tag.log 2>&1
# un-profiled wall times of the same commands
oracle/_ref/ref_driver bench /tmp/refk/f1080.f32 1920 1080 5 0.0 1.0 10.0 0.0 16384 5 40 > gpurun_out/ref_bench_extract_$tag.json 2>/dev/null
oracle/_ref/ref_driver benchmatch /tmp/refk/a.sift /tmp/refk/b.sift 3 20 > gpurun_out/ref_bench_match_$tag.json 2>/dev/null
