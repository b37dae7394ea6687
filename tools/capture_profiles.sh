#!/bin/bash
# Run ON THE GPU BOX (via gpurun): ncu launch list + one --set full capture per hot kernel -> gpurun_out/.
#   gpurun -- 'bash tools/capture_profiles.sh r01'
tag=${1:-r01}
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# launch list of 3 frames (the summariser keeps the last one)
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_$tag.csv python tools/run_frames.py 3 > gpurun_out/launches_$tag.log 2>&1
# launch list of the bench command itself (first 400 launches = its first ~57 frames; the process then runs on unprofiled)
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches_bench_$tag.csv python bench.py --steps 2 --warmup 3 --frames-per-step 64 --no-cpu-baseline > gpurun_out/launches_bench_$tag.log 2>&1
# per frame: blur_dog2<1> x4, blur_dog2<0> x1 (octave 0 is the first), find_points (all octaves), orient_desc
FULL="$NCU --set full --import-source on -f"
timeout 300 $FULL -k regex:k_blur_dog2 --launch-skip 10 --launch-count 1 -o gpurun_out/prof_${tag}_blur_dog_o0 python tools/run_frames.py 4 > /dev/null 2>&1
timeout 300 $FULL -k regex:k_find_points --launch-skip 2 --launch-count 1 -o gpurun_out/prof_${tag}_find_points python tools/run_frames.py 4 > /dev/null 2>&1
timeout 300 $FULL -k regex:k_orient_desc --launch-skip 2 --launch-count 1 -o gpurun_out/prof_${tag}_orient_desc python tools/run_frames.py 4 > /dev/null 2>&1
timeout 300 $FULL -k regex:k_match_tc --launch-skip 2 --launch-count 1 -o gpurun_out/prof_${tag}_match_tc python tools/run_match.py > /dev/null 2>&1
timeout 300 $FULL -k regex:k_rescore --launch-skip 2 --launch-count 1 -o gpurun_out/prof_${tag}_rescore python tools/run_match.py > /dev/null 2>&1
ls -la gpurun_out/
