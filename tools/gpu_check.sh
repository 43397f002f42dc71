#!/bin/bash
# One gpurun call: parity tests, bench (both arms), launch list, full ncu capture of each kernel.
# Usage (from the repo root on the GPU box): bash tools/gpu_check.sh <tag>
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,memory.total --format=csv > $out/gpu_$tag.txt; nproc >> $out/gpu_$tag.txt
python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest exit $?" | tee -a $out/pytest_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke exit $?" | tee -a $out/smoke_$tag.log
t0=$(date +%s); python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench exit $? ($(( $(date +%s) - t0 )) s)"
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err; echo "ref exit $?"
python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > $out/plain_bench_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > $out/ncu_launch_$tag.log 2>&1
# the headline launch itself (256 images): DRAM traffic for bench.py's roofline.traffic
ncu --set full --clock-control none --import-source on -k regex:blur -s 3 -c 1 -f -o $out/prof_bench_$tag \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > $out/ncu_bench_$tag.log 2>&1
ops="blur noise lowres lowres1080 lowresodd letterbox mixed jpeg"
python tools/profile_ops.py $ops > $out/plain_profile_$tag.log 2>&1 && \
ncu --set full --clock-control none -k regex:'blur|noise|lowres|letterbox|jpeg' -s 3 -c 30 -f \
    -o $out/prof_$tag python tools/profile_ops.py $ops > $out/ncu_full_$tag.log 2>&1
python tools/time_testset_driver.py 64 > $out/testset_driver_$tag.json 2> $out/testset_driver_$tag.err; echo "driver timing exit $?"
python tools/time_jpeg.py 64 > $out/time_jpeg_$tag.json 2> $out/time_jpeg_$tag.err; echo "jpeg timing exit $?"
python tools/time_jpegdec.py 128 > $out/time_jpegdec_$tag.json 2> $out/time_jpegdec_$tag.err; echo "jpeg decoder timing exit $?"
python tools/time_testset_sharded.py 548 gpu > $out/testset_548_$tag.json 2> $out/testset_548_$tag.err; echo "548-frame trees exit $?"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jpegdec --csv --log-file $out/jpegdec_launches_$tag.csv \
    python tools/time_jpegdec.py 128 > /dev/null 2>&1; echo "jpegdec launch list exit $?"
# gpurun merges back at most 64 MiB: drop the largest report rather than lose everything
while [ "$(du -sm $out | cut -f1)" -gt 60 ]; do
  big=$(ls -S $out/*.ncu-rep 2>/dev/null | head -1); [ -z "$big" ] && break
  echo "dropping $big ($(du -sm $big | cut -f1) MiB) to stay below the merge limit"; rm -f "$big"
done
echo "done"
