out=gpurun_out/r2f
mkdir -p $out
for dr in 8 4 16; do
DSK_DOWN_ROWS=$dr timeout 300 python bench.py --workload v2 --quant q2_k --n-layers 10 --steps 2 --warmup 2 --no-cpu-baseline --no-secondary --profile-token > $out/b_dr$dr.json 2> $out/b_dr$dr.err
grep -E "^down|^glu|token total" $out/b_dr$dr.err
python -c "import json;d=json.load(open('$out/b_dr$dr.json'));print('dr$dr value',d['value'])"
done
