# 2 GPUs: row-shard over NCCL (parity vs 1 GPU is covered by the gloo tests; here: does it run, what does it cost)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/c5_bench2.json 2> gpurun_out/c5_bench2.err; echo "bench2 rc=$?"
tail -5 gpurun_out/c5_bench2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c5_bench2.json'))
print('dp', d['ms_per_step'], d['value'])
print('strong', json.dumps(d.get('strong'))[:3000])
PY
# single GPU: dw flush variants
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/c5_bench1.json 2>gpurun_out/c5_b1.err
cd link-prediction-gnn_b200 && TWOWL_NVCC_DEFS=-DTWOWL_DW_FLUSH=2 python build.py --force > /dev/null 2>&1; cd ..
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/c5_bench1_f2.json 2>gpurun_out/c5_b1f2.err
cd link-prediction-gnn_b200 && TWOWL_NVCC_DEFS=-DTWOWL_DW_FLUSH=4 python build.py --force > /dev/null 2>&1; cd ..
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/c5_bench1_f4.json 2>gpurun_out/c5_b1f4.err
python - <<'PY'
import json
for f in ('c5_bench1','c5_bench1_f2','c5_bench1_f4'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); po=d['roofline']['per_op']
        print(f, round(d['ms_per_step'],2), 'pair_dw_gn', po['pair_dw_gn']['ms']/po['pair_dw_gn']['n'], 'seg_reduce', po['seg_reduce']['ms'], 'pair_conv', po['pair_conv']['ms'])
    except Exception as e: print(f, 'ERR', e)
PY
