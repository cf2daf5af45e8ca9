out=gpurun_out/final1
mkdir -p $out
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 > $out/bench_reference.json 2> $out/bench_reference.err ) 2> $out/time_ref.txt
tail -c 700 $out/bench_reference.json; tail -3 $out/time_ref.txt
( time timeout 1200 python bench.py --gpus 1 --steps 5 --warmup 3 > $out/bench_ours.json 2> $out/bench_ours.err ) 2> $out/time_ours.txt
tail -c 1500 $out/bench_ours.json; tail -3 $out/time_ours.txt
grep -a "bench +" $out/bench_ours.err | tail -25
