TAG=${TAG:-r02n}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
GFFM_MG_ROOT_FREE_MIN=2 MG_SELFTEST_TRANSPORTS=p2p_push,p2p_planes MG_SELFTEST_TIMING=0 timeout 300 $TR --master-port 29531 tools/mg_selftest.py > gpurun_out/${TAG}_selftest.out 2> gpurun_out/${TAG}_selftest.err; echo selftest rc=$?
grep "FAIL" gpurun_out/${TAG}_selftest.err | tail -5; tail -2 gpurun_out/${TAG}_selftest.out
GFFM_MG_ROOT_FREE_MIN=2 timeout 300 $TR --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-e2e --transport p2p_push,p2p_planes > gpurun_out/${TAG}_bench2_rootfree.json 2> gpurun_out/${TAG}_bench2_rootfree.err; echo rc=$?
python - <<'PY'
import json,glob,os
for f in sorted(glob.glob('gpurun_out/%s_bench2_*.json' % os.environ.get('TAG','r02n'))):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, round(d['ms_per_step'],3), 'gemm/step', r['gemm_ms_per_step_per_rank'], 'shards', d['config']['shards_match_local_product_on_all_ranks'], 'parity', d['parity_check'] and d['parity_check']['match'], d['clocks']['sm_mhz'], d['config']['multi_gpu_transport'], d['config']['warmup_trials_ms'], 'e2e', d['e2e'] and d['e2e'].get('ms_per_step'))
    except Exception as e: print(f, 'ERR', e)
PY
