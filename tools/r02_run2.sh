TAG=${TAG:-r02l}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
MG_SELFTEST_TIMING=0 timeout 300 $TR --master-port 29531 tools/mg_selftest.py > gpurun_out/${TAG}_selftest.out 2> gpurun_out/${TAG}_selftest.err; echo selftest rc=$?
grep "FAIL" gpurun_out/${TAG}_selftest.err | tail -5; tail -2 gpurun_out/${TAG}_selftest.out
timeout 300 $TR --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/${TAG}_bench2_tune.json 2> gpurun_out/${TAG}_bench2_tune.err; echo rc=$?
python - <<'PY'
import json,glob,os
for f in sorted(glob.glob('gpurun_out/%s_bench2_*.json' % os.environ.get('TAG','r02l'))):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, round(d['ms_per_step'],3), 'gemm/step', r['gemm_ms_per_step_per_rank'], 'shards', d['config']['shards_match_local_product_on_all_ranks'], 'parity', d['parity_check'] and d['parity_check']['match'], d['clocks']['sm_mhz'], d['config']['multi_gpu_transport'], d['config']['warmup_trials_ms'], 'e2e', d['e2e'] and d['e2e'].get('ms_per_step'))
    except Exception as e: print(f, 'ERR', e)
PY
