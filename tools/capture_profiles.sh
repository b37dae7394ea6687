#!/bin/bash
# Run ON THE GPU BOX (via gpurun): ncu launch list + one --set full capture per hot kernel -> gpurun_out/.
#   gpurun -- 'bash tools/capture_profiles.sh r02'      then, here:  python tools/summarize_ncu.py r02
tag=${1:-r02}
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# launch list of 3 frames (the summariser keeps the last one)
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_$tag.csv python tools/run_frames.py 3 > gpurun_out/launches_$tag.log 2>&1
# launch list of the bench command itself (first 400 launches; the process then runs on unprofiled)
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches_bench_$tag.csv python bench.py --steps 2 --warmup 3 --frames-per-step 64 --no-cpu-baseline --quick > gpurun_out/launches_bench_$tag.log 2>&1
# per frame: k_pyramid<1,0> (octave 0), k_down_chain<3>, k_pyramid<0,1> (octaves 1-4), k_find_points, k_orient_desc
FULL="$NCU --set full --import-source on -f"
# (k_pyramid launches alternate: octave 0, octaves 1-4, octave 0, ...)
timeout 300 $FULL -k regex:k_pyramid --launch-skip 4 --launch-count 1 -o gpurun_out/prof_${tag}_pyramid_o0 python tools/run_frames.py 4 > /dev/null 2>&1
timeout 300 $FULL -k regex:k_pyramid --launch-skip 5 --launch-count 1 -o gpurun_out/prof_${tag}_pyramid_rest python tools/run_frames.py 4 > /dev/null 2>&1
timeout 300 $FULL -k regex:k_down_chain --launch-skip 2 --launch-count 1 -o gpurun_out/prof_${tag}_down_chain python tools/run_frames.py 4 > /dev/null 2>&1
timeout 300 $FULL -k regex:k_find_points --launch-skip 2 --launch-count 1 -o gpurun_out/prof_${tag}_find_points python tools/run_frames.py 4 > /dev/null 2>&1
timeout 300 $FULL -k regex:k_orient_desc --launch-skip 2 --launch-count 1 -o gpurun_out/prof_${tag}_orient_desc python tools/run_frames.py 4 > /dev/null 2>&1
timeout 300 $FULL -k regex:k_match_tc --launch-skip 2 --launch-count 1 -o gpurun_out/prof_${tag}_match_tc python tools/run_match.py > /dev/null 2>&1
timeout 300 $FULL -k regex:k_rescore --launch-skip 2 --launch-count 1 -o gpurun_out/prof_${tag}_rescore python tools/run_match.py > /dev/null 2>&1
timeout 300 $FULL -k regex:k_hypotheses --launch-skip 1 --launch-count 1 -o gpurun_out/prof_${tag}_hypotheses python tools/allpairs_bench.py --sets 3 --points 2048 --size 640x480 --loops 1024 > /dev/null 2>&1
# all-pairs path: launch list + the kernels that only run there (cluster ImproveHomography, exact redo pass, fused RANSAC prep / score)
AP="python tools/allpairs_bench.py --sets 4 --points 8192 --improve 5 --reps 1"
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_allpairs_$tag.csv $AP > /dev/null 2>&1
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_match_$tag.csv python tools/run_match.py > /dev/null 2>&1
timeout 300 $FULL -k regex:k_improve_homography --launch-skip 2 --launch-count 1 -o gpurun_out/prof_${tag}_improve_homography $AP > /dev/null 2>&1
timeout 300 $FULL -k regex:k_match\< --launch-skip 2 --launch-count 1 -o gpurun_out/prof_${tag}_match_redo $AP > /dev/null 2>&1
timeout 300 $FULL -k regex:k_ransac_prep --launch-skip 2 --launch-count 1 -o gpurun_out/prof_${tag}_ransac_prep $AP > /dev/null 2>&1
timeout 300 $FULL -k regex:k_score --launch-skip 2 --launch-count 1 -o gpurun_out/prof_${tag}_score $AP > /dev/null 2>&1
ls -la gpurun_out/ | grep $tag
